"""ctypes binding of the CPU oracle (oracle/liborc.so).

TEST INFRASTRUCTURE ONLY - imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never imported by the product
package (go-dsp-guitar_b200/).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_double_p = C.POINTER(C.c_double)


def build(force=False):
    so = os.path.join(HERE, "liborc.so")
    srcs = [os.path.join(HERE, f) for f in ("orc_dsp.c", "orc_units.c", "orc_mix.c", "orc.h")]
    srcs += [os.path.join(HERE, "..", "include", f) for f in ("gdg_tables.h", "gdg_unit_params.def")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", HERE, "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(HERE, "liborc.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        vp, sz, u32, i32, dbl = C.c_void_p, C.c_size_t, C.c_uint32, C.c_int32, C.c_double
        sig = {
            "orc_fft_next_power_of_two": (C.c_int, [C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(u32)]),
            "orc_fft_fourier": (C.c_int, [vp, sz, C.c_int, C.c_int]),
            "orc_fft_inverse_fourier": (C.c_int, [vp, sz, C.c_int, C.c_int]),
            "orc_fft_real_fourier": (C.c_int, [vp, vp, sz, C.c_int]),
            "orc_fft_real_inverse_fourier": (C.c_int, [vp, vp, sz, C.c_int]),
            "orc_fft_shift": (None, [vp, sz, C.c_int]),
            "orc_resample_time_length": (sz, [sz, u32, u32]),
            "orc_resample_time": (None, [vp, sz, u32, u32, vp]),
            "orc_resample_frequency": (None, [vp, sz, u32, vp]),
            "orc_resample_oversample": (None, [vp, sz, vp, sz, u32]),
            "orc_prng_seed": (None, [vp, C.c_uint64]),
            "orc_prng_next_float": (dbl, [vp]),
            "orc_circular_create": (vp, [sz]),
            "orc_circular_destroy": (None, [vp]),
            "orc_circular_enqueue": (None, [vp, vp, sz]),
            "orc_circular_length": (sz, [vp]),
            "orc_circular_retrieve": (C.c_int, [vp, vp, sz]),
            "orc_wave_pcm16_decode": (None, [vp, sz, vp]),
            "orc_wave_pcm16_encode": (None, [vp, sz, vp]),
            "orc_wave_bytes_per_sample": (sz, [C.c_int]),
            "orc_wave_decode": (C.c_int, [C.c_int, vp, sz, vp]),
            "orc_wave_encode": (C.c_int, [C.c_int, vp, sz, vp]),
            "orc_filter_from_coefficients": (vp, [vp, sz, u32]),
            "orc_filter_destroy": (None, [vp]),
            "orc_filter_length": (sz, [vp]),
            "orc_filter_coefficients": (c_double_p, [vp]),
            "orc_filter_process": (C.c_int, [vp, vp, vp, sz]),
            "orc_filter_reduce": (vp, [vp, u32]),
            "orc_filter_normalize": (vp, [vp]),
            "orc_filter_multiply": (vp, [vp, dbl]),
            "orc_filter_set_compensation": (None, [vp, dbl]),
            "orc_irbank_create": (vp, []),
            "orc_irbank_destroy": (None, [vp]),
            "orc_irbank_add": (C.c_int, [vp, C.c_char_p, i32, vp, sz, u32]),
            "orc_irbank_add_raw": (C.c_int, [vp, C.c_char_p, i32, vp, sz, u32]),
            "orc_irbank_name_count": (sz, [vp]),
            "orc_irbank_name": (C.c_char_p, [vp, sz]),
            "orc_osd_create": (vp, [u32]),
            "orc_osd_destroy": (None, [vp]),
            "orc_osd_oversample": (C.c_int, [vp, vp, sz, vp, sz]),
            "orc_osd_decimate": (C.c_int, [vp, vp, sz, vp, sz]),
            "orc_chain_create": (vp, [vp]),
            "orc_chain_destroy": (None, [vp]),
            "orc_chain_append_unit": (C.c_int, [vp, C.c_int]),
            "orc_chain_remove_unit": (C.c_int, [vp, C.c_int]),
            "orc_chain_move_up": (C.c_int, [vp, C.c_int]),
            "orc_chain_move_down": (C.c_int, [vp, C.c_int]),
            "orc_chain_unit_type": (C.c_int, [vp, C.c_int]),
            "orc_chain_set_bypass": (C.c_int, [vp, C.c_int, C.c_int]),
            "orc_chain_get_bypass": (C.c_int, [vp, C.c_int, C.POINTER(C.c_int)]),
            "orc_chain_set_numeric": (C.c_int, [vp, C.c_int, C.c_char_p, i32]),
            "orc_chain_get_numeric": (C.c_int, [vp, C.c_int, C.c_char_p, C.POINTER(i32)]),
            "orc_chain_set_discrete": (C.c_int, [vp, C.c_int, C.c_char_p, C.c_char_p]),
            "orc_chain_get_discrete": (C.c_int, [vp, C.c_int, C.c_char_p, C.POINTER(C.c_char_p)]),
            "orc_chain_length": (C.c_int, [vp]),
            "orc_chain_process": (None, [vp, vp, vp, sz, u32]),
            "orc_last_error": (C.c_char_p, []),
            "orc_spatializer_create": (vp, [u32]),
            "orc_spatializer_destroy": (None, [vp]),
            "orc_spatializer_set_azimuth": (C.c_int, [vp, u32, dbl]),
            "orc_spatializer_set_distance": (C.c_int, [vp, u32, dbl]),
            "orc_spatializer_set_level": (C.c_int, [vp, u32, dbl]),
            "orc_spatializer_set_sample_rate": (None, [vp, u32]),
            "orc_spatializer_process": (None, [vp, vp, u32, vp, vp, u32, sz]),
            "orc_tuner_create": (vp, []),
            "orc_tuner_destroy": (None, [vp]),
            "orc_tuner_process": (None, [vp, vp, sz, u32]),
            "orc_tuner_analyze": (C.c_int, [vp, vp]),
            "orc_meter_create": (vp, [u32]),
            "orc_meter_destroy": (None, [vp]),
            "orc_meter_channel_count": (u32, [vp]),
            "orc_meter_enabled": (C.c_int, [vp]),
            "orc_meter_set_enabled": (None, [vp, C.c_int]),
            "orc_meter_process": (C.c_int, [vp, vp, vp, u32, u32]),
            "orc_meter_analyze": (C.c_int, [vp, u32, vp, vp]),
            "orc_meter_state": (C.c_int, [vp, u32, vp, vp, vp]),
            "orc_controller_process": (None, [vp, u32, vp, C.c_int, vp, vp, vp, u32, sz, u32]),
            "orc_bench_controller": (C.c_double, [vp, u32, vp, vp, sz, sz, u32, C.c_int, C.c_int]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f64(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def _c128(x):
    return np.ascontiguousarray(x, dtype=np.complex128)


class OracleError(RuntimeError):
    pass


# ---- fft ---------------------------------------------------------------------
SCALING_DEFAULT, SCALING_ORTHONORMAL, MODE_STANDARD, MODE_INPLACE = 0, 1, 2, 3


def next_power_of_two(v):
    p, e = C.c_uint64(), C.c_uint32()
    lib().orc_fft_next_power_of_two(v, C.byref(p), C.byref(e))
    return p.value, e.value


def fourier(vec, scaling=SCALING_DEFAULT, mode=MODE_INPLACE):
    v = _c128(vec).copy()
    if lib().orc_fft_fourier(_p(v), len(v), scaling, mode):
        raise OracleError("fourier")
    return v


def inverse_fourier(vec, scaling=SCALING_DEFAULT, mode=MODE_INPLACE):
    v = _c128(vec).copy()
    if lib().orc_fft_inverse_fourier(_p(v), len(v), scaling, mode):
        raise OracleError("inverse_fourier")
    return v


def real_fourier(x, scaling=SCALING_DEFAULT):
    x = _f64(x)
    out = np.zeros(len(x), dtype=np.complex128)
    if lib().orc_fft_real_fourier(_p(x), _p(out), len(x), scaling):
        raise OracleError("The number of elements in the vector must be even or one.")
    return out


def real_inverse_fourier(X, scaling=SCALING_DEFAULT):
    X = _c128(X).copy()
    out = np.zeros(len(X), dtype=np.float64)
    if lib().orc_fft_real_inverse_fourier(_p(X), _p(out), len(X), scaling):
        raise OracleError("The number of elements in the vector must be even or one.")
    return out


def shift(vec, inverse=False):
    v = _c128(vec).copy()
    lib().orc_fft_shift(_p(v), len(v), int(inverse))
    return v


# ---- resample ------------------------------------------------------------------
def resample_time(x, src, dst):
    x = _f64(x)
    n = lib().orc_resample_time_length(len(x), src, dst)
    out = np.zeros(n)
    lib().orc_resample_time(_p(x), len(x), src, dst, _p(out))
    return out


def resample_frequency(bins, n_target):
    b = _c128(bins)
    out = np.zeros(n_target, dtype=np.complex128)
    lib().orc_resample_frequency(_p(b), len(b), n_target, _p(out))
    return out


def resample_oversample(x, n_target, factor):
    x = _f64(x)
    out = np.zeros(n_target)
    lib().orc_resample_oversample(_p(x), len(x), _p(out), n_target, factor)
    return out


# ---- random ---------------------------------------------------------------------
class PRNG:
    def __init__(self, seed):
        self._s = (C.c_uint64 * 4)()
        lib().orc_prng_seed(C.addressof(self._s), seed & 0xFFFFFFFFFFFFFFFF)

    def next_float(self):
        return lib().orc_prng_next_float(C.addressof(self._s))

    def floats(self, n):
        return np.array([self.next_float() for _ in range(n)])


# ---- circular --------------------------------------------------------------------
class Circular:
    def __init__(self, size):
        self._h = lib().orc_circular_create(size)

    def __del__(self):
        lib().orc_circular_destroy(self._h)

    def enqueue(self, *elems):
        e = _f64(elems)
        lib().orc_circular_enqueue(self._h, _p(e), len(e))

    def length(self):
        return lib().orc_circular_length(self._h)

    def retrieve(self, n=None):
        n = self.length() if n is None else n
        out = np.zeros(n)
        if lib().orc_circular_retrieve(self._h, _p(out), n):
            raise OracleError("Target buffer must be of the same size as source buffer.")
        return out


# ---- wave --------------------------------------------------------------------------
def pcm16_decode(pcm):
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    out = np.zeros(len(pcm))
    lib().orc_wave_pcm16_decode(_p(pcm), len(pcm), _p(out))
    return out


def pcm16_encode(x):
    x = _f64(x)
    out = np.zeros(len(x), dtype=np.int16)
    lib().orc_wave_pcm16_encode(_p(x), len(x), _p(out))
    return out


WAVE_FORMATS = ["PCM8", "PCM16", "PCM24", "PCM32", "IEEE32", "IEEE64"]      # wave.go:737-838


def wave_bytes_per_sample(fmt):
    return lib().orc_wave_bytes_per_sample(WAVE_FORMATS.index(fmt) if isinstance(fmt, str) else fmt)


def wave_decode(fmt, data):
    f = WAVE_FORMATS.index(fmt) if isinstance(fmt, str) else fmt
    d = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8))
    n = len(d) // wave_bytes_per_sample(f)
    out = np.zeros(n)
    if lib().orc_wave_decode(f, _p(d), n, _p(out)):
        raise OracleError("unknown sample format")
    return out


def wave_encode(fmt, x):
    f = WAVE_FORMATS.index(fmt) if isinstance(fmt, str) else fmt
    x = _f64(x)
    out = np.zeros(len(x) * wave_bytes_per_sample(f), dtype=np.uint8)
    if lib().orc_wave_encode(f, _p(x), len(x), _p(out)):
        raise OracleError("unknown sample format")
    return out


# ---- filter ------------------------------------------------------------------------
class Filter:
    def __init__(self, coeffs=None, sample_rate=0, _handle=None):
        if _handle is not None:
            self._h = _handle
        else:
            c = _f64(coeffs)
            self._h = lib().orc_filter_from_coefficients(_p(c), len(c), sample_rate)

    def __del__(self):
        lib().orc_filter_destroy(self._h)

    def coefficients(self):
        n = lib().orc_filter_length(self._h)
        ptr = lib().orc_filter_coefficients(self._h)
        return np.ctypeslib.as_array(ptr, shape=(n,)).copy() if n else np.zeros(0)

    def process(self, x):
        x = _f64(x)
        out = np.zeros(len(x))
        r = lib().orc_filter_process(self._h, _p(x), _p(out), len(x))
        if r:
            raise OracleError("filter.Process failed (%d)" % r)
        return out

    def reduce(self, order):
        return Filter(_handle=lib().orc_filter_reduce(self._h, order))

    def normalize(self):
        return Filter(_handle=lib().orc_filter_normalize(self._h))

    def multiply(self, s):
        return Filter(_handle=lib().orc_filter_multiply(self._h, s))

    def set_compensation(self, g):
        lib().orc_filter_set_compensation(self._h, g)


class IRBank:
    def __init__(self):
        self._h = lib().orc_irbank_create()

    def __del__(self):
        lib().orc_irbank_destroy(self._h)

    def add(self, name, compensation_db, samples, source_rate):
        s = _f64(samples)
        lib().orc_irbank_add(self._h, name.encode(), compensation_db, _p(s), len(s), source_rate)

    def add_raw(self, name, compensation_db, samples, rate):
        s = _f64(samples)
        lib().orc_irbank_add_raw(self._h, name.encode(), compensation_db, _p(s), len(s), rate)

    def names(self):
        return [lib().orc_irbank_name(self._h, i).decode() for i in range(lib().orc_irbank_name_count(self._h))]


# ---- oversampling --------------------------------------------------------------------
class OversamplerDecimator:
    def __init__(self, factor):
        self.factor = factor
        self._h = lib().orc_osd_create(factor)
        if not self._h:
            raise OracleError("bad factor")

    def __del__(self):
        lib().orc_osd_destroy(self._h)

    def oversample(self, x):
        x = _f64(x)
        out = np.zeros(len(x) * self.factor)
        if lib().orc_osd_oversample(self._h, _p(x), len(x), _p(out), len(out)):
            raise OracleError("oversample")
        return out

    def decimate(self, x):
        x = _f64(x)
        out = np.zeros(len(x) // self.factor)
        if lib().orc_osd_decimate(self._h, _p(x), len(x), _p(out), len(out)):
            raise OracleError("decimate")
        return out


# ---- chain ----------------------------------------------------------------------------
UNIT_TYPES = ["signal_generator", "noise_gate", "bandpass", "auto_wah", "auto_yoy", "compressor", "octaver",
              "excess", "fuzz", "overdrive", "distortion", "tone_stack", "chorus", "flanger", "phaser", "tremolo",
              "ring_modulator", "delay", "reverb", "power_amp", "cabinet"]


class Chain:
    """Mirror of signal.Chain (signal/signal.go:21-36) over the oracle."""

    def __init__(self, bank=None):
        self._bank = bank
        self._h = lib().orc_chain_create(bank._h if bank is not None else None)

    def __del__(self):
        lib().orc_chain_destroy(self._h)

    @staticmethod
    def _err():
        return OracleError(lib().orc_last_error().decode())

    def append_unit(self, unit_type):
        if isinstance(unit_type, str):
            unit_type = UNIT_TYPES.index(unit_type)
        r = lib().orc_chain_append_unit(self._h, unit_type)
        if r < 0:
            raise self._err()
        return r

    def remove_unit(self, i):
        if lib().orc_chain_remove_unit(self._h, i):
            raise self._err()

    def move_up(self, i):
        if lib().orc_chain_move_up(self._h, i):
            raise self._err()

    def move_down(self, i):
        if lib().orc_chain_move_down(self._h, i):
            raise self._err()

    def unit_type(self, i):
        r = lib().orc_chain_unit_type(self._h, i)
        if r < 0:
            raise self._err()
        return r

    def set_bypass(self, i, b):
        if lib().orc_chain_set_bypass(self._h, i, int(b)):
            raise self._err()

    def get_bypass(self, i):
        v = C.c_int()
        if lib().orc_chain_get_bypass(self._h, i, C.byref(v)):
            raise self._err()
        return bool(v.value)

    def set_numeric_value(self, i, name, value):
        if lib().orc_chain_set_numeric(self._h, i, name.encode(), value):
            raise self._err()

    def get_numeric_value(self, i, name):
        v = C.c_int32()
        if lib().orc_chain_get_numeric(self._h, i, name.encode(), C.byref(v)):
            raise self._err()
        return v.value

    def set_discrete_value(self, i, name, value):
        if lib().orc_chain_set_discrete(self._h, i, name.encode(), value.encode()):
            raise self._err()

    def get_discrete_value(self, i, name):
        v = C.c_char_p()
        if lib().orc_chain_get_discrete(self._h, i, name.encode(), C.byref(v)):
            raise self._err()
        return v.value.decode()

    def length(self):
        return lib().orc_chain_length(self._h)

    def process(self, x, sample_rate):
        x = _f64(x)
        out = np.zeros(len(x))
        lib().orc_chain_process(self._h, _p(x), _p(out), len(x), sample_rate)
        return out


# ---- spatializer ------------------------------------------------------------------------
class Spatializer:
    def __init__(self, n):
        self.n = n
        self._h = lib().orc_spatializer_create(n)

    def __del__(self):
        lib().orc_spatializer_destroy(self._h)

    def set_azimuth(self, ch, v):
        if lib().orc_spatializer_set_azimuth(self._h, ch, v):
            raise OracleError("azimuth")

    def set_distance(self, ch, v):
        if lib().orc_spatializer_set_distance(self._h, ch, v):
            raise OracleError("distance")

    def set_level(self, ch, v):
        if lib().orc_spatializer_set_level(self._h, ch, v):
            raise OracleError("level")

    def set_sample_rate(self, rate):
        lib().orc_spatializer_set_sample_rate(self._h, rate)

    def process(self, inputs, aux=None):
        x = _f64(inputs)
        assert x.ndim == 2 and x.shape[0] == self.n
        n = x.shape[1]
        out = np.zeros((2, n))
        ip = (C.c_void_p * self.n)(*[x[i].ctypes.data for i in range(self.n)])
        op = (C.c_void_p * 2)(out[0].ctypes.data, out[1].ctypes.data)
        a = _f64(aux) if aux is not None else None
        lib().orc_spatializer_process(self._h, ip, self.n, _p(a) if a is not None else None, op, 2, n)
        return out


# ---- tuner ---------------------------------------------------------------------------------
class _TunerResult(C.Structure):
    _fields_ = [("cents", C.c_int8), ("frequency", C.c_double), ("note", C.c_char_p)]


class Tuner:
    def __init__(self):
        self._h = lib().orc_tuner_create()

    def __del__(self):
        lib().orc_tuner_destroy(self._h)

    def process(self, x, rate):
        x = _f64(x)
        lib().orc_tuner_process(self._h, _p(x), len(x), rate)

    def analyze(self):
        r = _TunerResult()
        if lib().orc_tuner_analyze(self._h, C.byref(r)):
            raise OracleError("analyze")
        return {"cents": r.cents, "frequency": r.frequency, "note": r.note.decode()}


# ---- level meter (level/level.go) --------------------------------------------------------
class Meter:
    def __init__(self, channels):
        self._h = lib().orc_meter_create(channels)
        self.count = channels

    def __del__(self):
        lib().orc_meter_destroy(self._h)

    @property
    def enabled(self):
        return bool(lib().orc_meter_enabled(self._h))

    def set_enabled(self, value):
        lib().orc_meter_set_enabled(self._h, 1 if value else 0)

    def process(self, buffers, rate):
        """buffers: one array (or None) per channel, lengths may differ."""
        rows = [None if b is None else _f64(b) for b in buffers]
        k = len(rows)
        bp = (C.c_void_p * max(k, 1))(*[None if r is None else r.ctypes.data for r in rows])
        ln = (C.c_size_t * max(k, 1))(*[0 if r is None else len(r) for r in rows])
        if lib().orc_meter_process(self._h, bp, ln, k, rate):
            raise OracleError("Number of input buffers (%d) does not match number of channels (%d) for this level meter." % (k, self.count))

    def analyze(self, channel):
        lv, pk = C.c_int32(), C.c_int32()
        if lib().orc_meter_analyze(self._h, channel, C.byref(lv), C.byref(pk)):
            raise OracleError("Requested analysis for channel %d, but level meter only has %d channels." % (channel, self.count))
        return lv.value, pk.value

    def state(self, channel):
        c, p, q = C.c_double(), C.c_double(), C.c_uint64()
        if lib().orc_meter_state(self._h, channel, C.byref(c), C.byref(p), C.byref(q)):
            raise OracleError("channel")
        return c.value, p.value, q.value


def controller_process(chains, tuner, tuner_channel, spat, inputs, sample_rate):
    """controller.process mirror: returns out[C (+2)][n]."""
    x = _f64(inputs)
    c, n = x.shape
    n_out = c + (2 if spat is not None else 0)
    out = np.zeros((n_out, n))
    ch = (C.c_void_p * c)(*[k._h for k in chains])
    ip = (C.c_void_p * c)(*[x[i].ctypes.data for i in range(c)])
    op = (C.c_void_p * n_out)(*[out[i].ctypes.data for i in range(n_out)])
    lib().orc_controller_process(ch, c, tuner._h if tuner else None, tuner_channel,
                                 spat._h if spat else None, ip, op, n_out, n, sample_rate)
    return out


def bench_controller(chains, spat, x, n, sample_rate, threads, pin=True):
    """orc_bench_controller: x [C][blocks * n] float64 through the chains (+ spatializer) on a pinned pthread pool;
    returns the wall time in seconds."""
    x = _f64(x)
    c, total = x.shape
    blocks = total // n
    ch = (C.c_void_p * c)(*[k._h for k in chains])
    return lib().orc_bench_controller(ch, c, spat._h if spat else None, _p(x), n, blocks, sample_rate, threads, 1 if pin else 0)
