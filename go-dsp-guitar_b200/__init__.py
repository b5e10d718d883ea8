"""go-dsp-guitar_b200 - host-side mirror of go-dsp-guitar's hot-path interfaces over libgdg.

The reference's host language (Go) is not available in this image, so the
reference-facing interface is mirrored here in Python over the C ABI
(include/gdg.h) with ctypes; the cgo binding a maintainer would add is in
INTEGRATION.md.  Names follow the reference:

    signal.Chain            -> Chain        (signal/signal.go:21-36)
    filter.ImpulseResponses -> IRBank       (filter/filter.go:79-82, Import :704)
    spatializer.Spatializer -> Spatializer  (spatializer/spatializer.go:30-41)
    tuner.Tuner             -> Tuner        (tuner/tuner.go:62-65)
    controller.process      -> controller_process (controller/controller.go:2648-2783)

Methods are the Go methods in snake_case with the same argument meaning; Go
`error` returns become GdgError carrying the reference's message text.

The package directory name contains a '-', so import it with
    importlib.import_module("go-dsp-guitar_b200")      (see gdg.py at the repo root)

PyTorch is not needed by this module; tests/bench use torch only for device
tensors, streams and torch.distributed.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIBRARY = os.path.join(HERE, "libgdg.so")      # the in-tree CUDA build; no environment override (no CPU path is reachable from here)

UNIT_TYPES = ["signal_generator", "noise_gate", "bandpass", "auto_wah", "auto_yoy", "compressor", "octaver",
              "excess", "fuzz", "overdrive", "distortion", "tone_stack", "chorus", "flanger", "phaser", "tremolo",
              "ring_modulator", "delay", "reverb", "power_amp", "cabinet"]          # effects.go:543-565
PARAMETER_TYPES = ["invalid", "discrete", "numeric"]                                  # effects.go:526-530


class GdgError(RuntimeError):
    pass


class _Parameter(C.Structure):
    _fields_ = [("name", C.c_char_p), ("type", C.c_int32), ("physical_unit", C.c_char_p),
                ("minimum", C.c_int32), ("maximum", C.c_int32), ("numeric_value", C.c_int32),
                ("discrete_value_index", C.c_int32), ("discrete_value_count", C.c_int32)]


class _TunerResult(C.Structure):
    _fields_ = [("cents", C.c_int8), ("frequency", C.c_double), ("note", C.c_char_p)]


_vp, _sz, _u32, _i32, _dbl, _int, _cp = C.c_void_p, C.c_size_t, C.c_uint32, C.c_int32, C.c_double, C.c_int, C.c_char_p

# every symbol include/gdg.h declares: name -> (restype, argtypes)
ABI = {
    "gdg_init": (_int, [_int]),
    "gdg_last_error": (_cp, []),
    "gdg_backend_name": (_cp, []),
    "gdg_version": (_cp, []),
    "gdg_launch_count": (C.c_uint64, []),
    "gdg_pinned_alloc": (_vp, [_sz]),
    "gdg_pinned_free": (None, [_vp]),
    "gdg_stream": (_vp, []),
    "gdg_synchronize": (_int, []),
    "gdg_unit_type_count": (_int, []),
    "gdg_unit_type_name": (_cp, [_int]),
    "gdg_irbank_import": (_vp, [_cp]),
    "gdg_irbank_create": (_vp, []),
    "gdg_irbank_destroy": (None, [_vp]),
    "gdg_irbank_add": (_int, [_vp, _cp, _i32, _vp, _sz, _u32]),
    "gdg_irbank_add_raw": (_int, [_vp, _cp, _i32, _vp, _sz, _u32]),
    "gdg_irbank_name_count": (_sz, [_vp]),
    "gdg_irbank_name": (_cp, [_vp, _sz]),
    "gdg_chain_create": (_vp, [_vp]),
    "gdg_chain_destroy": (None, [_vp]),
    "gdg_chain_append_unit": (_int, [_vp, _int]),
    "gdg_chain_remove_unit": (_int, [_vp, _int]),
    "gdg_chain_move_up": (_int, [_vp, _int]),
    "gdg_chain_move_down": (_int, [_vp, _int]),
    "gdg_chain_unit_type": (_int, [_vp, _int]),
    "gdg_chain_set_bypass": (_int, [_vp, _int, _int]),
    "gdg_chain_get_bypass": (_int, [_vp, _int, C.POINTER(_int)]),
    "gdg_chain_set_discrete_value": (_int, [_vp, _int, _cp, _cp]),
    "gdg_chain_get_discrete_value": (_int, [_vp, _int, _cp, C.POINTER(_cp)]),
    "gdg_chain_set_numeric_value": (_int, [_vp, _int, _cp, _i32]),
    "gdg_chain_get_numeric_value": (_int, [_vp, _int, _cp, C.POINTER(_i32)]),
    "gdg_chain_length": (_int, [_vp]),
    "gdg_chain_parameter_count": (_int, [_vp, _int]),
    "gdg_chain_parameter": (_int, [_vp, _int, _int, C.POINTER(_Parameter)]),
    "gdg_chain_parameter_discrete_value": (_cp, [_vp, _int, _int, _int]),
    "gdg_chain_process": (_int, [_vp, _vp, _vp, _sz, _u32]),
    "gdg_process_batch": (_int, [_vp, _sz, _vp, _vp, _sz, _u32]),
    "gdg_process_batch_device": (_int, [_vp, _sz, _vp, _vp, _sz, _sz, _u32, _vp]),
    "gdg_format_bytes_per_sample": (_sz, [_int]),
    "gdg_process_batch_encoded": (_int, [_vp, _sz, _vp, _int, _vp, _int, _sz, _u32]),
    "gdg_process_batch_device_encoded": (_int, [_vp, _sz, _vp, _int, _sz, _vp, _int, _sz, _sz, _u32, _vp]),
    "gdg_resample_time_length": (_sz, [_sz, _u32, _u32]),
    "gdg_resample_time": (C.c_long, [_vp, _sz, _u32, _u32, _vp, _sz]),
    "gdg_spatializer_create": (_vp, [_u32]),
    "gdg_spatializer_destroy": (None, [_vp]),
    "gdg_spatializer_set_azimuth": (_int, [_vp, _u32, _dbl]),
    "gdg_spatializer_get_azimuth": (_int, [_vp, _u32, C.POINTER(_dbl)]),
    "gdg_spatializer_set_distance": (_int, [_vp, _u32, _dbl]),
    "gdg_spatializer_get_distance": (_int, [_vp, _u32, C.POINTER(_dbl)]),
    "gdg_spatializer_set_level": (_int, [_vp, _u32, _dbl]),
    "gdg_spatializer_get_level": (_int, [_vp, _u32, C.POINTER(_dbl)]),
    "gdg_spatializer_input_count": (_u32, [_vp]),
    "gdg_spatializer_output_count": (_u32, [_vp]),
    "gdg_spatializer_set_sample_rate": (_int, [_vp, _u32]),
    "gdg_spatializer_process": (_int, [_vp, _vp, _u32, _vp, _vp, _u32, _sz]),
    "gdg_spatializer_process_device": (_int, [_vp, _vp, _sz, _vp, _vp, _sz, _sz, _int, _vp]),
    "gdg_tuner_create": (_vp, []),
    "gdg_tuner_destroy": (None, [_vp]),
    "gdg_tuner_process": (_int, [_vp, _vp, _sz, _u32]),
    "gdg_tuner_process_device": (_int, [_vp, _vp, _sz, _u32, _vp]),
    "gdg_tuner_analyze": (_int, [_vp, C.POINTER(_TunerResult)]),
    "gdg_meter_create": (_vp, [_u32, _vp]),
    "gdg_meter_destroy": (None, [_vp]),
    "gdg_meter_channel_count": (_u32, [_vp]),
    "gdg_meter_channel_name": (_cp, [_vp, _u32]),
    "gdg_meter_enabled": (_int, [_vp]),
    "gdg_meter_set_enabled": (_int, [_vp, _int]),
    "gdg_meter_process": (_int, [_vp, _vp, _u32, _sz, _u32]),
    "gdg_meter_process_device": (_int, [_vp, _u32, _vp, _sz, _u32, _sz, _u32, _vp]),
    "gdg_meter_analyze": (_int, [_vp, _u32, C.POINTER(_i32), C.POINTER(_i32)]),
    "gdg_meter_state": (_int, [_vp, _u32, C.POINTER(_dbl), C.POINTER(_dbl), C.POINTER(C.c_uint64)]),
    "gdg_controller_process": (_int, [_vp, _u32, _vp, _int, _vp, _vp, _vp, _u32, _sz, _u32]),
    "gdg_controller_process_metered": (_int, [_vp, _u32, _vp, _int, _vp, _vp, _vp, _vp, _u32, _sz, _u32]),
    "gdg_controller_submit": (_int, [_vp, _u32, _vp, _int, _vp, _vp, _vp, _vp, _u32, _sz, _u32]),
    "gdg_controller_wait": (_int, [_int]),
    "gdg_controller_submit_encoded": (_int, [_vp, _u32, _vp, _int, _vp, _vp, _vp, _int, _vp, _int, _u32, _sz, _u32]),
    "gdg_chain_filter_coefficients": (C.c_long, [_vp, _int, _vp, _sz]),
    "gdg_device_count": (_int, []),
    "gdg_use_device": (_int, [_int]),
    "gdg_current_device": (_int, []),
    "gdg_cluster_create": (_vp, [_vp, _int]),
    "gdg_cluster_destroy": (None, [_vp]),
    "gdg_cluster_size": (_int, [_vp]),
    "gdg_cluster_select": (_int, [_vp, _int]),
    "gdg_cluster_submit": (_int, [_vp, _vp, _u32, _vp, _int, _vp, _vp, _vp, _u32, _sz, _u32]),
    "gdg_cluster_wait": (_int, [_vp, _int]),
    "gdg_cluster_submit_aux": (_int, [_vp, _vp, _u32, _vp, _int, _vp, _vp, _vp, _vp, _int, _vp, _u32, _sz, _u32]),
    "gdg_cluster_submit_encoded": (_int, [_vp, _vp, _u32, _vp, _int, _vp, _vp, _vp, _int, _vp, _int, _u32, _sz, _u32]),
    "gdg_cluster_process": (_int, [_vp, _vp, _u32, _vp, _int, _vp, _vp, _vp, _u32, _sz, _u32]),
    "gdg_tuner_analyze_batch": (_int, [_vp, _u32, _vp]),
    "gdg_controller_process_aux": (_int, [_vp, _u32, _vp, _int, _vp, _vp, _vp, _vp, _int, _vp, _u32, _sz, _u32]),
    "gdg_controller_submit_aux": (_int, [_vp, _u32, _vp, _int, _vp, _vp, _vp, _vp, _int, _vp, _u32, _sz, _u32]),
    "gdg_process_batch_encoded_submit": (_int, [_vp, _sz, _vp, _int, _vp, _int, _sz, _u32]),
}

_LIBS = {}


def device_count(lib=None):
    return (lib if lib is not None else load()).gdg_device_count()


def use_device(device, lib=None):
    """objects created afterwards on this thread live on `device` (gdg_use_device)"""
    lib = lib if lib is not None else load()
    if lib.gdg_use_device(device) < 0:
        raise GdgError(lib.gdg_last_error().decode())


def load(path=None, _test_backend=False):
    """dlopen libgdg (default: the in-tree CUDA build) and bind the ABI.  Fails loudly if the
    library is missing - there is no Python / CPU fallback for any processing call.  Any library whose
    gdg_backend_name() is not the CUDA backend is REFUSED unless the caller is the test suite
    (tests/common.py passes _test_backend=True for tests/emu/libgdg_emu.so)."""
    path = os.path.abspath(path or DEFAULT_LIBRARY)
    if path not in _LIBS:
        if not os.path.exists(path):
            raise GdgError("libgdg not built: %s is missing (run `python go-dsp-guitar_b200/build.py`); "
                           "there is no CPU fallback" % path)
        lib = C.CDLL(path)
        for name, (res, args) in ABI.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        backend = lib.gdg_backend_name().decode()
        if not backend.startswith("cuda:") and not _test_backend:
            raise GdgError("%s is not the CUDA product library (backend '%s'): refused - there is no CPU path" % (path, backend))
        _LIBS[path] = lib
    return _LIBS[path]


def _f64(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class _Base:
    def __init__(self, lib=None):
        self._lib = lib if lib is not None else load()

    def _err(self):
        return GdgError(self._lib.gdg_last_error().decode("utf-8", "replace"))

    def _check(self, rc):
        if rc < 0:
            raise self._err()
        return rc


def init(device=-1, lib=None):
    lib = lib if lib is not None else load()
    if lib.gdg_init(device) != 0:
        raise GdgError(lib.gdg_last_error().decode())


def backend_name(lib=None):
    return (lib if lib is not None else load()).gdg_backend_name().decode()


def launch_count(lib=None):
    return (lib if lib is not None else load()).gdg_launch_count()


class IRBank(_Base):
    """filter.ImpulseResponses (filter/filter.go:79-82)."""

    def __init__(self, lib=None, _handle=None):
        super().__init__(lib)
        self._h = _handle if _handle is not None else self._lib.gdg_irbank_create()

    @classmethod
    def import_file(cls, index_json_path, lib=None):
        """filter.Import (filter/filter.go:704-802)."""
        lib = lib if lib is not None else load()
        h = lib.gdg_irbank_import(os.fsencode(index_json_path))
        if not h:
            raise GdgError(lib.gdg_last_error().decode())
        return cls(lib, _handle=h)

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.gdg_irbank_destroy(self._h)

    def add(self, name, compensation_db, samples, source_rate):
        s = _f64(samples)
        self._check(self._lib.gdg_irbank_add(self._h, name.encode(), compensation_db, _ptr(s), len(s), source_rate))

    def add_raw(self, name, compensation_db, samples, rate):
        s = _f64(samples)
        self._check(self._lib.gdg_irbank_add_raw(self._h, name.encode(), compensation_db, _ptr(s), len(s), rate))

    def names(self):
        return [self._lib.gdg_irbank_name(self._h, i).decode() for i in range(self._lib.gdg_irbank_name_count(self._h))]


class Chain(_Base):
    """signal.Chain (signal/signal.go:21-36) backed by the fused sm_100a segment kernel."""

    def __init__(self, bank=None, lib=None):
        super().__init__(lib if lib is not None else (bank._lib if bank is not None else None))
        self._bank = bank                                   # keep alive
        self._h = self._lib.gdg_chain_create(bank._h if bank is not None else None)
        if not self._h:
            raise self._err()

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.gdg_chain_destroy(self._h)

    def append_unit(self, unit_type):
        if isinstance(unit_type, str):
            unit_type = UNIT_TYPES.index(unit_type)
        r = self._lib.gdg_chain_append_unit(self._h, unit_type)
        if r < 0:
            raise self._err()
        return r

    def remove_unit(self, i):
        self._check(self._lib.gdg_chain_remove_unit(self._h, i))

    def move_up(self, i):
        self._check(self._lib.gdg_chain_move_up(self._h, i))

    def move_down(self, i):
        self._check(self._lib.gdg_chain_move_down(self._h, i))

    def unit_type(self, i):
        return self._check(self._lib.gdg_chain_unit_type(self._h, i))

    def set_bypass(self, i, bypass):
        self._check(self._lib.gdg_chain_set_bypass(self._h, i, int(bool(bypass))))

    def get_bypass(self, i):
        v = C.c_int()
        self._check(self._lib.gdg_chain_get_bypass(self._h, i, C.byref(v)))
        return bool(v.value)

    def set_discrete_value(self, i, name, value):
        self._check(self._lib.gdg_chain_set_discrete_value(self._h, i, name.encode(), value.encode()))

    def get_discrete_value(self, i, name):
        v = C.c_char_p()
        self._check(self._lib.gdg_chain_get_discrete_value(self._h, i, name.encode(), C.byref(v)))
        return v.value.decode()

    def set_numeric_value(self, i, name, value):
        self._check(self._lib.gdg_chain_set_numeric_value(self._h, i, name.encode(), int(value)))

    def get_numeric_value(self, i, name):
        v = C.c_int32()
        self._check(self._lib.gdg_chain_get_numeric_value(self._h, i, name.encode(), C.byref(v)))
        return v.value

    def parameters(self, i):
        """Chain.Parameters(id) -> list of effects.Parameter-like dicts (effects.go:69-78)."""
        n = self._check(self._lib.gdg_chain_parameter_count(self._h, i))
        out = []
        for k in range(n):
            p = _Parameter()
            self._check(self._lib.gdg_chain_parameter(self._h, i, k, C.byref(p)))
            vals = [self._lib.gdg_chain_parameter_discrete_value(self._h, i, k, j).decode()
                    for j in range(p.discrete_value_count)]
            out.append({"Name": p.name.decode(), "Type": p.type, "PhysicalUnit": p.physical_unit.decode(),
                        "Minimum": p.minimum, "Maximum": p.maximum, "NumericValue": p.numeric_value,
                        "DiscreteValueIndex": p.discrete_value_index, "DiscreteValues": vals})
        return out

    def length(self):
        return self._lib.gdg_chain_length(self._h)

    def filter_coefficients(self, i):
        """Filter.Coefficients() of the power amp in slot i (compiled on the device); empty before the first process()."""
        n = self._lib.gdg_chain_filter_coefficients(self._h, i, None, 0)
        if n < 0:
            raise self._err()
        out = np.zeros(n)
        if n and self._lib.gdg_chain_filter_coefficients(self._h, i, _ptr(out), n) < 0:
            raise self._err()
        return out

    def process(self, x, sample_rate):
        """Chain.Process(in, out, sampleRate) with host buffers (signal.go:361-414)."""
        x = _f64(x)
        out = np.empty_like(x)
        self._check(self._lib.gdg_chain_process(self._h, _ptr(x), _ptr(out), len(x), sample_rate))
        return out


def _handles(objs):
    return (C.c_void_p * len(objs))(*[o._h for o in objs])


def process_batch(chains, inputs, sample_rate):
    """The channel fan-out of controller.process (controller.go:2682-2705), host buffers [C][n]."""
    lib = chains[0]._lib
    x = _f64(inputs)
    c, n = x.shape
    out = np.empty_like(x)
    ip = (C.c_void_p * c)(*[x[i].ctypes.data for i in range(c)])
    op = (C.c_void_p * c)(*[out[i].ctypes.data for i in range(c)])
    if lib.gdg_process_batch(_handles(chains), c, ip, op, n, sample_rate) < 0:
        raise GdgError(lib.gdg_last_error().decode())
    return out


FORMATS = ["PCM8", "PCM16", "PCM24", "PCM32", "IEEE32", "IEEE64"]          # wave/wave.go:737-838


def process_batch_encoded(chains, rows, in_format, out_format, n, sample_rate):
    """gdg_process_batch_encoded: rows = list of bytes-like (one channel's payload each); returns list of bytes."""
    lib = chains[0]._lib
    fi, fo = FORMATS.index(in_format), FORMATS.index(out_format)
    ins = [np.frombuffer(bytes(r), dtype=np.uint8).copy() for r in rows]
    outs = [np.zeros(n * lib.gdg_format_bytes_per_sample(fo), dtype=np.uint8) for _ in rows]
    c = len(rows)
    ip = (C.c_void_p * c)(*[a.ctypes.data for a in ins])
    op = (C.c_void_p * c)(*[a.ctypes.data for a in outs])
    if lib.gdg_process_batch_encoded(_handles(chains), c, ip, fi, op, fo, n, sample_rate) < 0:
        raise GdgError(lib.gdg_last_error().decode())
    return [o.tobytes() for o in outs]


def resample_time(x, source_rate, target_rate, lib=None):
    """resample.Time (resample/resample.go:72-103) on the device."""
    lib = lib if lib is not None else load()
    x = _f64(x)
    n = lib.gdg_resample_time_length(len(x), source_rate, target_rate)
    out = np.zeros(n)
    r = lib.gdg_resample_time(_ptr(x), len(x), source_rate, target_rate, _ptr(out), n)
    if r < 0:
        raise GdgError(lib.gdg_last_error().decode())
    return out[:r]


class BatchRunner:
    """Pre-marshalled handle array for repeated device-resident batches (bench / streaming)."""

    def __init__(self, chains):
        self.chains = list(chains)
        self._lib = chains[0]._lib
        self._arr = _handles(self.chains)

    def process_device(self, d_in_ptr, d_out_ptr, stride, n, sample_rate, stream=None):
        if self._lib.gdg_process_batch_device(self._arr, len(self.chains), d_in_ptr, d_out_ptr, stride, n, sample_rate, stream) < 0:
            raise GdgError(self._lib.gdg_last_error().decode())

    def process_device_encoded(self, d_in_ptr, in_format, in_stride_bytes, d_out_ptr, out_format, out_stride_bytes, n, sample_rate, stream=None):
        """gdg_process_batch_device_encoded: payload rows already on the device, decoded / encoded inside the segment kernels"""
        if self._lib.gdg_process_batch_device_encoded(self._arr, len(self.chains), d_in_ptr, FORMATS.index(in_format), in_stride_bytes,
                                                      d_out_ptr, FORMATS.index(out_format), out_stride_bytes, n, sample_rate, stream) < 0:
            raise GdgError(self._lib.gdg_last_error().decode())

    def process_host(self, in_ptrs, out_ptrs, n, sample_rate):
        if self._lib.gdg_process_batch(self._arr, len(self.chains), in_ptrs, out_ptrs, n, sample_rate) < 0:
            raise GdgError(self._lib.gdg_last_error().decode())


class Spatializer(_Base):
    """spatializer.Spatializer (spatializer/spatializer.go:30-41)."""

    def __init__(self, input_channels, lib=None):
        super().__init__(lib)
        self.n = input_channels
        self._h = self._lib.gdg_spatializer_create(input_channels)
        if not self._h:
            raise self._err()

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.gdg_spatializer_destroy(self._h)

    def set_azimuth(self, ch, v):
        self._check(self._lib.gdg_spatializer_set_azimuth(self._h, ch, v))

    def get_azimuth(self, ch):
        v = C.c_double()
        self._check(self._lib.gdg_spatializer_get_azimuth(self._h, ch, C.byref(v)))
        return v.value

    def set_distance(self, ch, v):
        self._check(self._lib.gdg_spatializer_set_distance(self._h, ch, v))

    def get_distance(self, ch):
        v = C.c_double()
        self._check(self._lib.gdg_spatializer_get_distance(self._h, ch, C.byref(v)))
        return v.value

    def set_level(self, ch, v):
        self._check(self._lib.gdg_spatializer_set_level(self._h, ch, v))

    def get_level(self, ch):
        v = C.c_double()
        self._check(self._lib.gdg_spatializer_get_level(self._h, ch, C.byref(v)))
        return v.value

    def get_input_count(self):
        return self._lib.gdg_spatializer_input_count(self._h)

    def get_output_count(self):
        return self._lib.gdg_spatializer_output_count(self._h)

    def set_sample_rate(self, rate):
        self._check(self._lib.gdg_spatializer_set_sample_rate(self._h, rate))

    def process(self, inputs, aux=None):
        x = _f64(inputs)
        c, n = x.shape
        out = np.zeros((2, n))
        ip = (C.c_void_p * c)(*[x[i].ctypes.data for i in range(c)])
        op = (C.c_void_p * 2)(out[0].ctypes.data, out[1].ctypes.data)
        a = _f64(aux) if aux is not None else None
        self._check(self._lib.gdg_spatializer_process(self._h, ip, c, _ptr(a) if a is not None else None, op, 2, n))
        return out

    def process_device(self, d_in_ptr, stride, d_out_ptr, out_stride, n, d_aux_ptr=None, accumulate=False, stream=None):
        self._check(self._lib.gdg_spatializer_process_device(self._h, d_in_ptr, stride, d_aux_ptr, d_out_ptr, out_stride,
                                                             n, int(accumulate), stream))


class Tuner(_Base):
    """tuner.Tuner (tuner/tuner.go:62-65)."""

    def __init__(self, lib=None):
        super().__init__(lib)
        self._h = self._lib.gdg_tuner_create()
        if not self._h:
            raise self._err()

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.gdg_tuner_destroy(self._h)

    def process(self, samples, sample_rate):
        s = _f64(samples)
        self._check(self._lib.gdg_tuner_process(self._h, _ptr(s), len(s), sample_rate))

    def process_device(self, d_ptr, n, sample_rate, stream=None):
        self._check(self._lib.gdg_tuner_process_device(self._h, d_ptr, n, sample_rate, stream))

    def analyze(self):
        r = _TunerResult()
        self._check(self._lib.gdg_tuner_analyze(self._h, C.byref(r)))
        return {"cents": r.cents, "frequency": r.frequency, "note": r.note.decode()}


def tuner_analyze_batch(tuners):
    """gdg_tuner_analyze_batch: Analyze for several tuners with ONE batch of device launches."""
    if not tuners:
        return []
    lib = tuners[0]._lib
    res = (_TunerResult * len(tuners))()
    if lib.gdg_tuner_analyze_batch(_handles(tuners), len(tuners), res) < 0:
        raise GdgError(lib.gdg_last_error().decode())
    return [{"cents": r.cents, "frequency": r.frequency, "note": r.note.decode()} for r in res]


class Meter(_Base):
    """level.Meter (level/level.go:57-65)."""

    def __init__(self, channels, names=None, lib=None):
        super().__init__(lib)
        if names is not None and len(names) != channels:     # level.go:325-326
            raise GdgError("Failed to create channel meter. Requested channel meter for %d channels, but provided %d channel names."
                           % (channels, len(names)))
        arr = None
        if names is not None:
            self._names = [nm.encode() for nm in names]
            arr = (C.c_char_p * max(channels, 1))(*self._names)
        self._h = self._lib.gdg_meter_create(channels, arr)
        if not self._h:
            raise self._err()

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.gdg_meter_destroy(self._h)

    def channel_count(self):
        return self._lib.gdg_meter_channel_count(self._h)

    def channel_name(self, channel):
        r = self._lib.gdg_meter_channel_name(self._h, channel)
        if r is None:
            raise self._err()
        return r.decode()

    def enabled(self):
        return bool(self._lib.gdg_meter_enabled(self._h))

    def set_enabled(self, value):
        self._check(self._lib.gdg_meter_set_enabled(self._h, 1 if value else 0))

    def process(self, buffers, sample_rate):
        """buffers: one equally long array (or None) per channel."""
        rows = [None if b is None else _f64(b) for b in buffers]
        k = len(rows)
        n = max([len(r) for r in rows if r is not None], default=0)
        if any(r is not None and len(r) != n for r in rows):
            raise GdgError("gdg_meter_process: buffers of one call must have the same length")
        bp = (C.c_void_p * max(k, 1))(*[None if r is None else r.ctypes.data for r in rows])
        self._check(self._lib.gdg_meter_process(self._h, bp, k, n, sample_rate))

    def process_device(self, first_channel, d_rows_ptr, stride, rows, n, sample_rate, stream=None):
        self._check(self._lib.gdg_meter_process_device(self._h, first_channel, d_rows_ptr, stride, rows, n, sample_rate, stream))

    def analyze(self, channel):
        lv, pk = _i32(), _i32()
        self._check(self._lib.gdg_meter_analyze(self._h, channel, C.byref(lv), C.byref(pk)))
        return lv.value, pk.value

    def state(self, channel):
        c, p, q = _dbl(), _dbl(), C.c_uint64()
        self._check(self._lib.gdg_meter_state(self._h, channel, C.byref(c), C.byref(p), C.byref(q)))
        return c.value, p.value, q.value


def controller_process(chains, tuner, tuner_channel, spat, inputs, sample_rate, meter=None, metronome=None, metronome_to_master=True):
    """controller.process (controller.go:2648-2783) with host buffers: returns out[C (+2)][n].  `metronome`: the row the
    metronome wrote for this block (controller.go:2724-2738), mixed into the master when metronome_to_master."""
    lib = chains[0]._lib
    x = _f64(inputs)
    c, n = x.shape
    n_out = c + (2 if spat is not None else 0)
    out = np.zeros((n_out, n))
    ip = (C.c_void_p * c)(*[x[i].ctypes.data for i in range(c)])
    op = (C.c_void_p * n_out)(*[out[i].ctypes.data for i in range(n_out)])
    if metronome is not None:
        m = _f64(metronome)
        rc = lib.gdg_controller_process_aux(_handles(chains), c, tuner._h if tuner else None, tuner_channel, spat._h if spat else None,
                                            meter._h if meter else None, ip, _ptr(m), 1 if metronome_to_master else 0, op, n_out, n, sample_rate)
    else:
        rc = lib.gdg_controller_process_metered(_handles(chains), c, tuner._h if tuner else None, tuner_channel,
                                                spat._h if spat else None, meter._h if meter else None, ip, op, n_out, n, sample_rate)
    if rc < 0:
        raise GdgError(lib.gdg_last_error().decode())
    return out


def controller_submit(chains, tuner, tuner_channel, spat, inputs, outputs, sample_rate, meter=None, metronome=None, metronome_to_master=True):
    """gdg_controller_submit: enqueue one block without waiting.  `inputs` [C][n] and `outputs` [C (+2)][n] are caller-owned
    C-contiguous float64 arrays that must stay alive (inputs untouched) until controller_wait(ticket).  Returns the ticket."""
    lib = chains[0]._lib
    if inputs.dtype != np.float64 or outputs.dtype != np.float64 or not inputs.flags.c_contiguous or not outputs.flags.c_contiguous:
        raise GdgError("controller_submit needs C-contiguous float64 arrays")
    c, n = inputs.shape
    n_out = outputs.shape[0]
    ip = (C.c_void_p * c)(*[inputs[i].ctypes.data for i in range(c)])
    op = (C.c_void_p * n_out)(*[outputs[i].ctypes.data for i in range(n_out)])
    if metronome is not None:                     # the row is uploaded inside the call: it need not outlive it
        t = lib.gdg_controller_submit_aux(_handles(chains), c, tuner._h if tuner else None, tuner_channel, spat._h if spat else None,
                                          meter._h if meter else None, ip, _ptr(_f64(metronome)), 1 if metronome_to_master else 0, op, n_out, n, sample_rate)
    else:
        t = lib.gdg_controller_submit(_handles(chains), c, tuner._h if tuner else None, tuner_channel, spat._h if spat else None,
                                      meter._h if meter else None, ip, op, n_out, n, sample_rate)
    if t < 0:
        raise GdgError(lib.gdg_last_error().decode())
    return t


def controller_wait(ticket, lib=None):
    lib = lib if lib is not None else load()
    if lib.gdg_controller_wait(ticket) < 0:
        raise GdgError(lib.gdg_last_error().decode())


class Cluster(_Base):
    """gdg_cluster: several GPUs driven by ONE process (controller.go:3339-3341 is one process with a worker per channel).
    Create the members' objects after select(member); process() = controller.process over all members."""

    def __init__(self, devices, lib=None):
        super().__init__(lib)
        arr = (C.c_int * len(devices))(*devices)
        self._h = self._lib.gdg_cluster_create(arr, len(devices))
        if not self._h:
            raise self._err()
        self.devices = list(devices)

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.gdg_cluster_destroy(self._h)

    def select(self, member):
        self._check(self._lib.gdg_cluster_select(self._h, member))

    def process(self, chains, tuner, tuner_channel, spats, inputs, sample_rate, meters=None, metronome=None, metronome_to_master=True):
        """controller.process over the members: returns out[C (+2)][n].  meters: one Meter per member (2 * cnt_m + 3 channels,
        entries may be None); metronome: the click row of this block (see controller_process)."""
        x = _f64(inputs)
        c, n = x.shape
        n_out = c + (2 if spats else 0)
        out = np.zeros((n_out, n))
        ip = (C.c_void_p * c)(*[x[i].ctypes.data for i in range(c)])
        op = (C.c_void_p * n_out)(*[out[i].ctypes.data for i in range(n_out)])
        sp = _handles(spats) if spats else None
        if meters is None and metronome is None:
            self._check(self._lib.gdg_cluster_process(self._h, _handles(chains), c, tuner._h if tuner else None, tuner_channel, sp, ip, op, n_out, n, sample_rate))
            return out
        mt = (C.c_void_p * len(self.devices))(*[(m._h if m is not None else None) for m in meters]) if meters else None
        click = _f64(metronome) if metronome is not None else None
        t = self._lib.gdg_cluster_submit_aux(self._h, _handles(chains), c, tuner._h if tuner else None, tuner_channel, sp, mt, ip,
                                             _ptr(click) if click is not None else None, 1 if metronome_to_master else 0, op, n_out, n, sample_rate)
        self._check(t)
        self._check(self._lib.gdg_cluster_wait(self._h, t))
        return out

    def submit(self, chains, tuner, tuner_channel, spats, inputs, outputs, sample_rate, meters=None, metronome=None, metronome_to_master=True):
        """gdg_cluster_submit(_aux): enqueue one block on every member without waiting; `inputs` [C][n] / `outputs` [C (+2)][n]
        are caller-owned C-contiguous float64 arrays that stay alive until wait(ticket).  Returns the ticket."""
        if inputs.dtype != np.float64 or outputs.dtype != np.float64 or not inputs.flags.c_contiguous or not outputs.flags.c_contiguous:
            raise GdgError("Cluster.submit needs C-contiguous float64 arrays")
        c, n = inputs.shape
        n_out = outputs.shape[0]
        ip = (C.c_void_p * c)(*[inputs[i].ctypes.data for i in range(c)])
        op = (C.c_void_p * n_out)(*[outputs[i].ctypes.data for i in range(n_out)])
        sp = _handles(spats) if spats else None
        mt = (C.c_void_p * len(self.devices))(*[(m._h if m is not None else None) for m in meters]) if meters else None
        click = _f64(metronome) if metronome is not None else None
        t = self._lib.gdg_cluster_submit_aux(self._h, _handles(chains), c, tuner._h if tuner else None, tuner_channel, sp, mt, ip,
                                             _ptr(click) if click is not None else None, 1 if metronome_to_master else 0, op, n_out, n, sample_rate)
        self._check(t)
        return t

    def submit_encoded(self, chains, tuner, tuner_channel, spats, rows, in_format, outputs, out_format, n, sample_rate, meters=None):
        """gdg_cluster_submit_encoded: rows / outputs are lists of caller-owned uint8 arrays (one payload row per channel, the
        master pair last) that stay alive until wait(ticket)."""
        fi, fo = FORMATS.index(in_format), FORMATS.index(out_format)
        c, n_out = len(rows), len(outputs)
        ip = (C.c_void_p * c)(*[a.ctypes.data for a in rows])
        op = (C.c_void_p * n_out)(*[a.ctypes.data for a in outputs])
        sp = _handles(spats) if spats else None
        mt = (C.c_void_p * len(self.devices))(*[(m._h if m is not None else None) for m in meters]) if meters else None
        t = self._lib.gdg_cluster_submit_encoded(self._h, _handles(chains), c, tuner._h if tuner else None, tuner_channel, sp, mt, ip, fi, op, fo,
                                                 n_out, n, sample_rate)
        self._check(t)
        return t

    def wait(self, ticket):
        self._check(self._lib.gdg_cluster_wait(self._h, ticket))
