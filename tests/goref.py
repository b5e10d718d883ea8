"""goref.py - a SECOND, independent restatement of the reference's hot path in Python.

TEST INFRASTRUCTURE ONLY.  Written directly from the Go sources under /root/reference
(file:line cited per function) - NOT from oracle/orc_*.c - so that a misreading of
a unit in the C oracle (oracle/) shows up as a difference between the two
(tests/test_oracle_crosscheck.py diffs them at <= 1e-12).  Structure follows the Go code
one statement at a time: Python floats are IEEE binary64 like Go's float64, `math`
is the platform libm (Go's math.* are not correctly rounded either: <= 1-2 ulp apart).
Sequential per-sample loops, so use small blocks.

Deliberately different tools from the C oracle where the result is mathematically
the same: filter.Process uses numpy's FFT (pocketfft) instead of a radix-2 restatement
(fft/fft.go), IR resampling uses vectorised numpy.

Go untyped constant expressions (`0.001 * math.Pi`, `2.0 / math.Pi`, ...) are exact
rational arithmetic rounded ONCE to float64 (Go spec, "Constant expressions"); they
are reproduced with fractions.Fraction below.
"""
import math
from fractions import Fraction

import numpy as np

# math.Pi as the Go source spells it (go/src/math/const.go), an exact decimal constant
_PI = Fraction("3.14159265358979323846264338327950288419716939937510582097494459")


def _c(fr):
    return float(fr)                      # Fraction -> nearest float64 (correctly rounded)


# effects/effects.go:49-55
MATH_DEGREE_TO_RADIANS = _c(_PI / 180)
MATH_PI_THOUSANDTH = _c(Fraction(1, 1000) * _PI)
MATH_QUARTER_PI = _c(Fraction(1, 4) * _PI)
MATH_TWO_OVER_PI = _c(2 / _PI)
MATH_TWO_PI = _c(2 * _PI)
MATH_TWO_PI_FIFTH = _c(Fraction(4, 10) * _PI)
MATH_TWO_PI_HUNDREDTH = _c(Fraction(2, 100) * _PI)
MATH_PI = _c(_PI)
NUM_FILTERS = 8                           # effects/effects.go:62

U32_MAX = 0xFFFFFFFF


def decibels_to_factor(db):               # effects/effects.go:389-394
    return math.pow(10.0, 0.05 * float(db))


def factor_to_decibels(f):                # effects/effects.go:399-402 (math.Log10(0) = -Inf, of a negative = NaN)
    if f == 0.0:
        return -math.inf
    if f < 0.0 or f != f:
        return math.nan
    return 20.0 * math.log10(f)


def sign_float(x):                        # effects/effects.go:424-438
    return -1.0 if x < 0.0 else (1.0 if x > 0.0 else 0.0)


def clamp1(x):
    return -1.0 if x < -1.0 else (1.0 if x > 1.0 else x)


def go_mod(x, y):
    """math.Mod: result has the sign of x (C fmod)."""
    return math.fmod(x, y)


def u32(x):
    """uint32(float64) conversion for in-range non-negative values (Go truncates toward zero)."""
    return int(x) & U32_MAX


class Unit:
    """effects.unitStruct (effects/effects.go:96-384): a parameter store; numeric values are int32, discrete are strings."""
    TYPE = None
    NUMERIC = {}                          # name -> (min, max, default)
    DISCRETE = {}                         # name -> (values, default index)

    def __init__(self):
        self.num = {k: v[2] for k, v in self.NUMERIC.items()}
        self.disc = {k: v[0][v[1]] for k, v in self.DISCRETE.items()}

    def set_numeric(self, name, value):
        mn, mx, _ = self.NUMERIC[name]
        if value < mn or value > mx:
            raise ValueError("out of range")
        self.num[name] = int(value)

    def set_discrete(self, name, value):
        if value not in self.DISCRETE[name][0]:
            raise ValueError("invalid discrete value")
        self.disc[name] = value

    def follower(self, follow, envelope, sample_abs, d_inv, d):
        """The envelope follower shared by compressor/autowah/autoyoy/octaver/fuzz (e.g. compressor.go:44-58)."""
        if follow == "envelope":
            envelope *= d_inv
            if sample_abs > envelope:
                envelope = sample_abs
        elif follow == "level":
            diff = sample_abs - envelope
            envelope += diff * d
        else:
            envelope = 1.0
        return envelope


# ------------------------------------------------------------------------------------------------
class NoiseGate(Unit):                    # effects/noisegate.go:19-96; parameters :101-135
    NUMERIC = {"threshold_open": (-60, 0, -20), "threshold_close": (-60, 0, -40), "hold_time": (0, 1000, 50)}

    def __init__(self):
        super().__init__()
        self.gate_open = False
        self.on_hold_since = 0

    def process(self, x, fs):
        level_open, level_close, hold_time = self.num["threshold_open"], self.num["threshold_close"], self.num["hold_time"]
        fac_open, fac_close = decibels_to_factor(level_open), decibels_to_factor(level_close)
        out = [0.0] * len(x)
        if level_open < level_close:
            out = list(x)
            self.gate_open = True
            self.on_hold_since = 0
            return out
        hold_samples = u32(math.floor(((0.001 * float(hold_time)) * float(fs)) + 0.5))
        gate_open, on_hold_since = self.gate_open, self.on_hold_since
        for i, s in enumerate(x):
            amp = abs(s)
            if amp > fac_open:
                gate_open = True
            if amp > fac_close:
                on_hold_since = 0
            if on_hold_since >= hold_samples:
                gate_open = False
            fac = 1.0 if gate_open else 0.0
            out[i] = fac * s
            if on_hold_since < U32_MAX:
                on_hold_since += 1
        self.gate_open, self.on_hold_since = gate_open, on_hold_since
        return out


class Bandpass(Unit):                     # effects/bandpass.go:20-98
    NUMERIC = {"frequency_1": (1, 20000, 300), "frequency_2": (1, 20000, 3000)}
    DISCRETE = {"filter_order": (["2", "4", "6", "8"], 0)}

    def __init__(self):
        super().__init__()
        self.hcv, self.lcv = [], []

    def process(self, x, fs):
        half = int(self.disc["filter_order"]) >> 1
        fa, fb = self.num["frequency_1"], self.num["frequency_2"]
        if fa > fb:
            fa, fb = fb, fa
        if len(self.hcv) != half:
            self.hcv = [0.0] * half
        if len(self.lcv) != half:
            self.lcv = [0.0] * half
        m2pi = -MATH_TWO_PI / float(fs)
        k_hp = 1.0 - math.exp(m2pi * float(fa))
        k_lp = 1.0 - math.exp(m2pi * float(fb))
        out = [0.0] * len(x)
        for i, s in enumerate(x):
            pre = s
            for j in range(half):
                hcv = self.hcv[j]
                diff = pre - hcv
                hcv += diff * k_hp
                self.hcv[j] = hcv
                lcv = self.lcv[j]
                diff -= lcv
                iv = lcv
                lcv += diff * k_lp
                self.lcv[j] = lcv
                pre = clamp1(iv)
            out[i] = pre
        return out


class AutoWah(Unit):                      # effects/autowah.go:20-130
    NUMERIC = {"level_1": (-60, 0, -40), "level_2": (-60, 0, -10), "frequency_1": (1, 20000, 300), "frequency_2": (1, 20000, 6000)}
    DISCRETE = {"follow": (["envelope", "level"], 1)}

    def __init__(self):
        super().__init__()
        self.envelope = 0.0
        self.hcvs = [0.0] * NUM_FILTERS
        self.lcvs = [0.0] * NUM_FILTERS

    def process(self, x, fs):
        follow = self.disc["follow"]
        la, lb, fa, fb = self.num["level_1"], self.num["level_2"], self.num["frequency_1"], self.num["frequency_2"]
        if la > lb:
            la, lb = lb, la
            fa, fb = fb, fa
        la, lb, fa, fb = float(la), float(lb), float(fa), float(fb)
        with np.errstate(all="ignore"):
            slope = float(np.float64(fb - fa) / np.float64(lb - la))      # Go float division: x/0 = +-Inf or NaN, no panic
        sr = float(fs)
        d_inv = math.exp(-20.0 / sr)
        d = 1.0 - d_inv
        envelope = self.envelope
        gain_comp = math.pow(2.0, NUM_FILTERS)
        out = [0.0] * len(x)
        for i, s in enumerate(x):
            envelope = self.follower(follow, envelope, abs(s), d_inv, d)
            level = factor_to_decibels(envelope)
            if level <= la:
                freq = fa
            elif level >= lb:
                freq = fb
            else:
                freq = fa + (slope * (level - la))
            k = 1.0 - math.exp(-freq / sr)
            lcv = s
            for j in range(NUM_FILTERS):
                hcv = self.hcvs[j]
                diff = lcv - hcv
                hcv += diff * k
                self.hcvs[j] = hcv
                lcv = self.lcvs[j]
                diff -= lcv
                lcv += diff * k
                self.lcvs[j] = lcv
            out[i] = clamp1(gain_comp * lcv)
        self.envelope = envelope
        return out


class _GoIndexPanic(IndexError):
    """The reference would panic here (index out of range)."""


def _delayed(x, buf, idx):
    if idx >= 0:
        return x[idx]
    p = len(buf) + idx
    if p < 0:
        raise _GoIndexPanic("buffer[%d]: the reference panics (index out of range)" % p)
    return buf[p]


def _fractional_delay(x, buf, i, delay_samples):
    """The two-tap read shared by autoyoy/chorus/flanger/phaser (e.g. autoyoy.go:116-141): weights
    1-(d-floor d) and 1-(ceil d-d); indices below zero come from the history buffer (oldest first)."""
    early, late = math.floor(delay_samples), math.ceil(delay_samples)
    se = _delayed(x, buf, i - int(early))
    sl = _delayed(x, buf, i - int(late))
    we = 1.0 - (delay_samples - early)
    wl = 1.0 - (late - delay_samples)
    return (we * se) + (wl * sl)


def _push_history(buf, x):
    """copy(buffer[0:boundary], buffer[n:size]); copy(buffer[boundary:], in)  or  copy(buffer, in[-boundary:]) (e.g. autoyoy.go:146-155)"""
    size, n = len(buf), len(x)
    boundary = size - n
    if boundary >= 0:
        buf[0:boundary] = buf[n:size]
        buf[boundary:size] = list(x)
    else:
        buf[:] = list(x[-boundary:n])


class AutoYoy(Unit):                      # effects/autoyoy.go:19-156
    NUMERIC = {"level_1": (-60, 0, -40), "level_2": (-60, 0, -10), "depth": (0, 100, 100)}
    DISCRETE = {"follow": (["envelope", "level"], 1)}

    def __init__(self):
        super().__init__()
        self.envelope = 0.0
        self.buffer = []

    def process(self, x, fs):
        follow = self.disc["follow"]
        la, lb = self.num["level_1"], self.num["level_2"]
        depth_a, depth_b = 0.0, 0.01 * float(self.num["depth"])
        if la > lb:
            la, lb = lb, la
            depth_a, depth_b = depth_b, depth_a
        la, lb = float(la), float(lb)
        with np.errstate(all="ignore"):
            slope = float(np.float64(depth_b - depth_a) / np.float64(lb - la))
        sr = float(fs)
        sr_inv = 1.0 / sr
        d_inv = math.exp(-20.0 * sr_inv)
        d = 1.0 - d_inv
        max_delay = int(math.floor((0.01 * sr) + 0.5))
        if len(self.buffer) != max_delay:
            self.buffer = [0.0] * max_delay
        buf = self.buffer
        envelope = self.envelope
        out = [0.0] * len(x)
        for i, s in enumerate(x):
            envelope = self.follower(follow, envelope, abs(s), d_inv, d)
            level = factor_to_decibels(envelope)
            if level <= la:
                dfac = depth_a
            elif level >= lb:
                dfac = depth_b
            else:
                dfac = depth_a + (slope * (level - la))
            dsmp = (0.01 * dfac) * sr
            out[i] = (0.5 * s) + (0.5 * _fractional_delay(x, buf, i, dsmp))
        self.envelope = envelope
        _push_history(buf, x)
        return out


class Compressor(Unit):                   # effects/compressor.go:18-84
    NUMERIC = {"gain_limit": (0, 30, 30), "target_level": (-30, 0, -20)}
    DISCRETE = {"follow": (["envelope", "level"], 1)}

    def __init__(self):
        super().__init__()
        self.envelope = 0.0

    def process(self, x, fs):
        follow = self.disc["follow"]
        limit = decibels_to_factor(self.num["gain_limit"])
        target = decibels_to_factor(self.num["target_level"])
        d_inv = math.exp(-20.0 / float(fs))
        d = 1.0 - d_inv
        envelope = self.envelope
        out = [0.0] * len(x)
        for i, s in enumerate(x):
            envelope = self.follower(follow, envelope, abs(s), d_inv, d)
            with np.errstate(all="ignore"):
                gain = float(np.float64(target) / np.float64(envelope))
            if gain > limit:
                gain = limit
            out[i] = clamp1(gain * s)
        self.envelope = envelope
        return out


class Octaver(Unit):                      # effects/octaver.go:21-139
    NUMERIC = {"level_octave_up": (-60, 0, -20), "level_clean": (-60, 0, -20), "level_dist": (-60, 0, -20),
               "level_octave_down_first": (-60, 0, -20), "level_octave_down_second": (-60, 0, -20), "level_hysteresis": (-60, 0, -20)}
    DISCRETE = {"follow": (["envelope", "level"], 1)}

    def __init__(self):
        super().__init__()
        self.previous_polarity = 0.0
        self.octave_register = 0
        self.envelope = 0.0
        self.ccv = 0.0

    def process(self, x, fs):
        follow = self.disc["follow"]
        f_up = decibels_to_factor(self.num["level_octave_up"])
        f_clean = decibels_to_factor(self.num["level_clean"])
        f_dist = decibels_to_factor(self.num["level_dist"])
        f_d1 = decibels_to_factor(self.num["level_octave_down_first"])
        f_d2 = decibels_to_factor(self.num["level_octave_down_second"])
        f_hyst = decibels_to_factor(self.num["level_hysteresis"])
        pol, reg, envelope, ccv = self.previous_polarity, self.octave_register, self.envelope, self.ccv
        d_inv = math.exp(-20.0 / float(fs))
        d = 1.0 - d_inv
        out = [0.0] * len(x)
        for i, s in enumerate(x):
            sa = abs(s)
            envelope = self.follower(follow, envelope, sa, d_inv, d)
            square = s * s
            sign = sign_float(s)
            hyst = envelope * f_hyst
            if sign != 0.0 and sign != pol and sa > hyst:
                reg = (reg + 1) & 0x7
                pol = sign
            first_down = -1.0 if (reg & 0x2) != 0 else 1.0
            second_down = -1.0 if (reg & 0x4) != 0 else 1.0
            pre = f_clean * s
            if envelope > 0.0001:
                pre += f_up * (square / envelope)
            pre += f_dist * (sign * envelope)
            pre += f_d1 * (first_down * envelope)
            pre += f_d2 * (second_down * envelope)
            ccv += (pre - ccv) * d
            pre -= ccv
            out[i] = clamp1(pre)
        self.previous_polarity, self.octave_register, self.envelope, self.ccv = pol, reg, envelope, ccv
        return out


# ---- random.PRNG (random/random.go:23-55) -------------------------------------------------------
class PRNG:
    A, M = 16807, 2147483647

    def __init__(self, seed):
        self.x = ((64979 * (seed & 0xFFFFFFFFFFFFFFFF)) + 83) % (1 << 64) % self.M

    def next_float(self):
        self.x = (self.A * self.x) % self.M
        return float(self.x) / float(self.M - 1)


class SignalGenerator(Unit):              # effects/signalgenerator.go:20-153
    NUMERIC = {"input_amplitude": (0, 100, 100), "input_gain": (-60, 0, 0), "signal_frequency": (1, 20000, 440),
               "signal_amplitude": (0, 100, 100), "signal_gain": (-60, 0, 0)}
    DISCRETE = {"signal_type": (["sine", "triangle", "square", "sawtooth", "noise"], 0)}

    def __init__(self):
        super().__init__()
        self.phase = 0.0
        self.prng = None

    def process(self, x, fs):
        fac_in = (0.01 * float(self.num["input_amplitude"])) * decibels_to_factor(self.num["input_gain"])
        fac_sig = (0.01 * float(self.num["signal_amplitude"])) * decibels_to_factor(self.num["signal_gain"])
        phase = self.phase
        inc = MATH_TWO_PI * (float(self.num["signal_frequency"]) / float(fs))
        two_over_pi = _c(2 / _PI)
        n = len(x)
        kind = self.disc["signal_type"]
        out = [0.0] * n
        if kind == "noise":
            if self.prng is None:
                self.prng = PRNG(1337)
            for i, s in enumerate(x):
                r = self.prng.next_float()
                out[i] = (fac_in * s) + (fac_sig * (1.0 - (2.0 * r)))
        else:
            for i, s in enumerate(x):
                cp = go_mod(phase + (float(i) * inc), MATH_TWO_PI)
                if kind == "sine":
                    sig = math.sin(cp)
                elif kind == "triangle":
                    sig = (two_over_pi * cp) - 1.0 if cp < MATH_PI else 3.0 - (two_over_pi * cp)
                elif kind == "square":
                    sig = sign_float(MATH_PI - cp)
                else:
                    sig = cp / MATH_PI
                    if cp > MATH_PI:
                        sig -= 2.0
                out[i] = (fac_in * s) + (fac_sig * sig)
            phase += float(n) * inc
            phase = go_mod(phase, MATH_TWO_PI)
        self.phase = phase
        return out


# ---- resample (resample/resample.go) ------------------------------------------------------------
def lanczos_kernel(x, a):                 # resample.go:10-31
    if x == 0:
        return 1.0
    if -a < x < a:
        pix = MATH_PI * x
        pixa = pix / a
        return (a * (math.sin(pix) * math.sin(pixa))) / (pix * pix)
    return 0.0


def lanczos_interpolate(s, x, a):         # resample.go:36-66
    idx_inc = int(math.floor(x)) + 1
    n = len(s)
    total = 0.0
    for i in range(idx_inc - a, idx_inc + a):
        if 0 <= i < n:
            total += s[i] * lanczos_kernel(x - float(i), float(a))
    return total


def resample_time(samples, source_rate, target_rate):      # resample.go:72-103
    lf = float(len(samples)) * (float(target_rate) / float(source_rate))
    fl = math.floor(lf)
    length = int(fl)
    if fl == lf:
        length -= 1
    dx = float(source_rate) / float(target_rate)
    return [lanczos_interpolate(samples, float(i) * dx, 3) for i in range(length)]


def resample_frequency(bins, num_target):                   # resample.go:109-142
    re = [float(b.real) for b in bins]
    im = [float(b.imag) for b in bins]
    dx = float(len(bins)) / float(num_target)
    return [complex(lanczos_interpolate(re, float(i) * dx, 3), lanczos_interpolate(im, float(i) * dx, 3)) for i in range(num_target)]


def resample_oversample(source, target_len, factor):        # resample.go:148-176
    dx = 1.0 / float(factor)
    out = [0.0] * target_len
    for i in range(target_len):
        if i % factor == 0:
            out[i] = source[i // factor]
        else:
            out[i] = lanczos_interpolate(source, float(i) * dx, 3)
    return out


# ---- fft (fft/fft.go): only what the hot path needs, through numpy ---------------------------------
def next_power_of_two(v):                 # fft.go:393-409
    p = 1
    while p < v:
        p <<= 1
    return p


def real_fourier(x):                      # fft.go:744-856, SCALING_DEFAULT: the full n-point spectrum of a real sequence
    return np.fft.fft(np.asarray(x, dtype=np.float64))


def real_inverse_fourier(spec):           # fft.go:863-990: symmetrise (899-906), real parts of DC / Nyquist (908-913, 947-951), 1/n
    spec = np.asarray(spec, dtype=np.complex128)
    n = len(spec)
    h = n // 2
    half = np.empty(h + 1, dtype=np.complex128)
    half[0] = spec[0].real
    half[h] = spec[h].real
    if h > 1:
        half[1:h] = 0.5 * (spec[1:h] + np.conj(spec[n - 1:h:-1]))
    return np.fft.irfft(half, n)


# ---- filter (filter/filter.go) --------------------------------------------------------------------
SAMPLE_RATES = [22050, 32000, 44100, 48000, 88200, 96000, 192000]     # filter.go:25-33


class Filter:
    """filter.filterStruct: coefficients + lazily built spectrum + overlap tail (filter.go:58-68)."""

    def __init__(self, coeffs, rate, compensation=0.0):
        self.coeffs = [float(c) for c in coeffs]
        self.rate = rate
        self.comp = compensation
        self.filter_complex = None        # len == fftSize when valid
        self.tail = np.zeros(0)

    def add(self, other):                 # filter.go:167-253
        if other is None:
            return self
        if self.rate != other.rate:
            raise ValueError("Cannot add filters: Sample rates do not match.")
        n = max(len(self.coeffs), len(other.coeffs))
        res = list(self.coeffs) + [0.0] * (n - len(self.coeffs))
        for i, c in enumerate(other.coeffs):
            res[i] += c
        return Filter(res, self.rate, 0.0)

    def multiply(self, scalar):           # filter.go:270-323 (gainCompensation of the result is 0.0)
        return Filter([scalar * c for c in self.coeffs], self.rate, 0.0)

    def normalize(self):                  # filter.go:328-336 with estimateGain :127-138
        total = 0.0
        for c in self.coeffs:
            total += c * c
        gain = math.sqrt(total)
        with np.errstate(all="ignore"):
            fac = float(np.float64(self.comp) / np.float64(gain))
        return self.multiply(fac)

    def reduce(self, order):              # filter.go:520-604
        n = len(self.coeffs)
        n_src, n_tgt = next_power_of_two(n), next_power_of_two(order)
        if n <= order:
            return self
        padded = list(self.coeffs) + [0.0] * (n_src - n)
        fr = real_fourier(padded)
        pos_src = (n_src >> 1) + 1
        half_tgt = n_tgt >> 1
        pos_new = resample_frequency(list(fr[:pos_src]), half_tgt + 1)
        fr_new = np.zeros(n_tgt, dtype=np.complex128)
        fr_new[:len(pos_new)] = pos_new[:n_tgt]          # copy(frNew, frPosNew) stops at the shorter slice
        for i in range(1, half_tgt):
            fr_new[n_tgt - i] = np.conj(pos_new[i])
        target = real_inverse_fourier(fr_new)
        return Filter(list(target[:order]), self.rate, self.comp)

    def process(self, x):                 # filter.go:342-515
        n = len(x)
        L = len(self.coeffs)
        out = np.zeros(n)
        if L == 0:
            return out
        x = np.asarray(x, dtype=np.float64)
        n_power, block = next_power_of_two(n), next_power_of_two(L)
        num_blocks = n_power // block
        if n_power % block != 0:
            num_blocks += 1
        fft_size = block << 1
        for i in range(num_blocks):
            if self.filter_complex is None or len(self.filter_complex) != fft_size:
                padded = np.zeros(fft_size)
                padded[:L] = self.coeffs
                self.filter_complex = real_fourier(padded)
            if len(self.tail) != fft_size:
                self.tail = np.zeros(fft_size)
            lb = i * block
            ub = min(lb + block, n)
            ns = ub - lb
            buf = np.zeros(fft_size)
            buf[:ns] = x[lb:ub]
            y = real_inverse_fourier(real_fourier(buf) * self.filter_complex)
            pre = y + self.tail
            out[lb:ub] = np.clip(pre[:ns], -1.0, 1.0)
            self.tail[:fft_size - ns] = pre[ns:]
            self.tail[fft_size - ns:] = 0.0
        return out


class ImpulseResponses:                   # filter.impulseResponsesStruct
    def __init__(self):
        self.responses = []               # (name, rate, gainCompensation, data)

    def import_one(self, name, compensation_db, samples, source_rate):      # one descriptor of filter.Import, filter.go:736-790
        fac = math.pow(10.0, 0.05 * float(compensation_db))
        for rate in SAMPLE_RATES:
            self.responses.append((name, rate, fac, resample_time(list(samples), source_rate, rate)))

    add = import_one                      # name used by the oracle's Python wrapper

    def add_raw(self, name, compensation_db, samples, rate):
        self.responses.append((name, rate, math.pow(10.0, 0.05 * float(compensation_db)), [float(s) for s in samples]))

    def create_filter(self, name, rate):  # filter.go:619-660
        for n, r, comp, data in self.responses:
            if n == name and r == rate:
                return Filter(data, r, comp)
        return None

    def names(self):                      # filter.go:681-699
        out = []
        for n, _r, _c, _d in self.responses:
            if n not in out:
                out.append(n)
        return out


# ---- oversampling (oversampling/oversampling.go) -----------------------------------------------------
ATTENUATION_HALF_DECIBEL = 0.9440608762859234       # oversampling.go:13
LOOKAHEAD_ONE_SIDE = 4
LOOKAHEAD_BOTH_SIDES = 2 * LOOKAHEAD_ONE_SIDE
_AA = {}


def aa_coefficients(factor):
    """The anti-aliasing FIRs of oversampling.go:239-317 (x2, 77 taps) / :357-513 (x4, 155 taps), extracted from the Go
    literals by tests/golden/extract_goref_tables.py into tests/golden/goref_tables.json."""
    if not _AA:
        import json
        import os
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "goref_tables.json")) as f:
            t = json.load(f)
        _AA[2], _AA[4] = t["aa2"], t["aa4"]
    return _AA[factor]


class OversamplerDecimator:               # oversampling.go:40-184
    def __init__(self, factor):
        self.factor = factor
        self.flt = Filter(aa_coefficients(factor), 0) if factor > 1 else None
        self.pre = []

    def oversample(self, x):              # oversampling.go:54-115
        f = self.factor
        if f <= 1:
            return list(x)
        n = len(x)
        pre_size = n + LOOKAHEAD_BOTH_SIDES
        if len(self.pre) != pre_size:
            self.pre = [0.0] * pre_size
        pre = self.pre
        pre[0:LOOKAHEAD_BOTH_SIDES] = pre[pre_size - LOOKAHEAD_BOTH_SIDES:pre_size]
        pre[LOOKAHEAD_BOTH_SIDES:pre_size] = list(x)
        post = resample_oversample(pre, ((pre_size - 1) * f) + 1, f)
        start = LOOKAHEAD_ONE_SIDE * f
        return post[start:start + n * f]

    def decimate(self, x, n_out):         # oversampling.go:126-184
        f = self.factor
        if f <= 1:
            return list(x)
        buf = self.flt.process(x)
        return [ATTENUATION_HALF_DECIBEL * float(buf[f * i]) if f * i < len(buf) else 0.0 for i in range(n_out)]


class OversampledUnit(Unit):
    """The Oversample -> processOversampled -> Decimate wrapper shared by excess / fuzz / overdrive / distortion
    (e.g. overdrive.go:83-144); each unit owns one x2 and one x4 oversampler (overdrive.go:150-151)."""

    def __init__(self):
        super().__init__()
        self.os2, self.os4 = OversamplerDecimator(2), OversamplerDecimator(4)

    def process(self, x, fs):
        o = self.disc["oversampling"]
        factor = 2 if o == "2" else (4 if o == "4" else 1)
        if factor > 1:
            osd = self.os4 if factor == 4 else self.os2
            up = osd.oversample(x)
            shaped = self.process_oversampled(up, (factor * fs) & U32_MAX)
            return osd.decimate(shaped, len(x))
        return self.process_oversampled(x, fs)


_OS = (["- NONE -", "2", "4"], 0)


class Excess(OversampledUnit):            # effects/excess.go:22-66
    NUMERIC = {"gain": (-30, 30, 0), "level": (-30, 0, 0)}
    DISCRETE = {"oversampling": _OS}

    def process_oversampled(self, x, fs):
        gain, level = decibels_to_factor(self.num["gain"]), decibels_to_factor(self.num["level"])
        out = [0.0] * len(x)
        for i, s in enumerate(x):
            pre = gain * s
            ap = abs(pre)
            exceeded = ap > 1.0
            negative = pre < 0.0
            section = int(0.5 * math.floor(ap + 1.0))          # int32(): truncation toward zero
            odd = (section % 2) != 0
            inverted = odd != (exceeded and negative)
            ex = go_mod(ap + 1.0, 2.0)
            if exceeded:
                pre = 1.0 - ex if inverted else ex - 1.0
            out[i] = level * pre
        return out


class Distortion(OversampledUnit):        # effects/distortion.go:21-49
    NUMERIC = {"boost": (0, 30, 0), "gain": (-30, 30, 0), "level": (-30, 0, 0)}
    DISCRETE = {"oversampling": _OS}

    def process_oversampled(self, x, fs):
        gain = decibels_to_factor(self.num["boost"] + self.num["gain"])
        level = decibels_to_factor(self.num["level"])
        return [level * clamp1(gain * s) for s in x]


class Overdrive(OversampledUnit):         # effects/overdrive.go:28-78
    NUMERIC = {"boost": (0, 30, 0), "gain": (-30, 30, 0), "drive": (0, 100, 100), "level": (-30, 0, 0)}
    DISCRETE = {"valve": (["ECC82 (12AU7)", "ECC83 (12AX7)"], 1), "oversampling": _OS}

    def process_oversampled(self, x, fs):
        gain = decibels_to_factor(self.num["boost"] + self.num["gain"])
        drive = 0.01 * float(self.num["drive"])
        clean = 1.0 - drive
        level = decibels_to_factor(self.num["level"])
        valve = self.disc["valve"]
        out = [0.0] * len(x)
        for i, s in enumerate(x):
            arg = gain * s
            dist = 0.0
            if valve == "ECC82 (12AU7)":
                dist = MATH_TWO_OVER_PI * math.atan(MATH_QUARTER_PI * arg)
            elif valve == "ECC83 (12AX7)":
                try:
                    e = math.exp(-arg)
                except OverflowError:
                    e = math.inf
                dist = (2.0 / (1.0 + e)) - 1.0
            out[i] = level * ((drive * dist) + (clean * s))
        return out


class Fuzz(OversampledUnit):              # effects/fuzz.go:24-108
    NUMERIC = {"bias": (-100, 100, 50), "boost": (0, 30, 0), "gain": (-30, 30, 0), "fuzz": (0, 100, 100), "level": (-30, 0, 0)}
    DISCRETE = {"follow": (["envelope", "level"], 1), "oversampling": _OS}

    def __init__(self):
        super().__init__()
        self.envelope = 0.0
        self.ccv = 0.0

    def process_oversampled(self, x, fs):
        follow = self.disc["follow"]
        bias = 0.01 * float(self.num["bias"])
        gain = decibels_to_factor(self.num["boost"] + self.num["gain"])
        ff = 0.01 * float(self.num["fuzz"])
        ffi = 1.0 - ff
        level = decibels_to_factor(self.num["level"])
        envelope, ccv = self.envelope, self.ccv
        d_inv = math.exp(-20.0 / float(fs))
        d = 1.0 - d_inv
        out = [0.0] * len(x)
        for i, s in enumerate(x):
            envelope = self.follower(follow, envelope, abs(s), d_inv, d)
            pre = clamp1(gain * (s - (bias * envelope)))
            pre = (ff * pre) + (ffi * s)
            diff = pre - ccv
            ccv += diff * d
            pre -= ccv
            out[i] = level * clamp1(pre)
        self.envelope, self.ccv = envelope, ccv
        return out


class ToneStack(Unit):                    # effects/tonestack.go:19-100
    NUMERIC = {"low": (-30, 0, 0), "middle": (-30, 0, -2), "presence": (-30, 0, -5), "high": (-30, 0, -5)}

    def __init__(self):
        super().__init__()
        self.hcv, self.lcv = [], []

    def process(self, x, fs):
        freqs = [20.0, 300.0, 3000.0, 6000.0, 20000.0]
        facs = [decibels_to_factor(self.num[k]) for k in ("low", "middle", "presence", "high")]
        nb = 4
        if len(self.hcv) != nb:
            self.hcv = [0.0] * nb
        if len(self.lcv) != nb:
            self.lcv = [0.0] * nb
        m2pi = -MATH_TWO_PI / float(fs)
        out = [0.0] * len(x)
        for i, s in enumerate(x):
            total = 0.0
            for j in range(nb):
                hcv, lcv = self.hcv[j], self.lcv[j]
                k_hp = 1.0 - math.exp(m2pi * freqs[j])
                k_lp = 1.0 - math.exp(m2pi * freqs[j + 1])
                diff = s - hcv
                hcv += diff * k_hp
                diff -= lcv
                pre = lcv
                lcv += diff * k_lp
                self.hcv[j], self.lcv[j] = hcv, lcv
                total += facs[j] * pre
            out[i] = clamp1(total)
        return out


class Cabinet(Unit):                      # effects/cabinet.go:27-162
    DISCRETE = {"type": (["- DEFAULT -"], 0)}

    def __init__(self):
        super().__init__()
        self.hcv = [0.0] * 3
        self.lcv = [0.0] * 4

    def process(self, x, fs):
        buf = list(x)
        m2pi = -MATH_TWO_PI / float(fs)
        for i, f in enumerate([300.0, 120.0, 80.0]):
            hcv = self.hcv[i]
            k = 1.0 - math.exp(m2pi * f)
            for j, s in enumerate(buf):
                diff = s - hcv
                buf[j] = diff
                hcv += diff * k
            self.hcv[i] = hcv
        for i, f in enumerate([3000.0, 4000.0, 5000.0, 6000.0]):
            lcv = self.lcv[i]
            k = 1.0 - math.exp(m2pi * f)
            for j, s in enumerate(buf):
                diff = s - lcv
                buf[j] = lcv
                lcv += diff * k
            self.lcv[i] = lcv
        return [clamp1(s) for s in buf]


class Chorus(Unit):                       # effects/chorus.go:19-130
    NUMERIC = {"depth": (0, 100, 100), "speed": (1, 100, 30)}

    def __init__(self):
        super().__init__()
        self.buffer = []
        self.previous_phase = 0.0

    def process(self, x, fs):
        depth = 0.1 * float(self.num["depth"])
        depth = 0.0 if depth < 0.0 else (10.0 if depth > 10.0 else depth)
        ang = MATH_PI_THOUSANDTH * float(self.num["speed"])
        sr = float(fs)
        max_delay = int(math.floor((0.05 * sr) + 0.5))
        prev = self.previous_phase
        if len(self.buffer) != max_delay:
            self.buffer = [0.0] * max_delay
        buf = self.buffer
        out = [0.0] * len(x)
        for i, s in enumerate(x):
            t = float(i) / sr
            zero_phase = go_mod(prev + (ang * t), MATH_TWO_PI)
            eff = 0.0
            for j in range(5):
                phase = go_mod(zero_phase + (MATH_TWO_PI_FIFTH * float(j)), MATH_TWO_PI)
                offset = depth * math.sin(phase)
                dsmp = (0.001 * (40.0 + offset)) * sr
                eff += 0.2 * _fractional_delay(x, buf, i, dsmp)
            out[i] = (0.5 * s) + (0.5 * eff)
        self.previous_phase = go_mod(prev + (ang * (float(len(buf)) / sr)), MATH_TWO_PI)
        _push_history(buf, x)
        return out


class Flanger(Unit):                      # effects/flanger.go:19-117
    NUMERIC = {"depth": (0, 100, 100), "speed": (1, 100, 10)}

    def __init__(self):
        super().__init__()
        self.buffer = []
        self.previous_phase = 0.0

    def weights(self):
        return 0.5, 0.5

    def process(self, x, fs):
        depth = 0.01 * float(self.num["depth"])
        depth = 0.0 if depth < 0.0 else (1.0 if depth > 1.0 else depth)
        ang = MATH_TWO_PI_HUNDREDTH * float(self.num["speed"])
        wa, wb = self.weights()
        sr = float(fs)
        sr_inv = 1.0 / sr
        max_delay = int(math.floor((0.002 * sr) + 0.5))
        prev = self.previous_phase
        if len(self.buffer) != max_delay:
            self.buffer = [0.0] * max_delay
        buf = self.buffer
        out = [0.0] * len(x)
        for i, s in enumerate(x):
            t = float(i) * sr_inv
            phase = go_mod(prev + (ang * t), MATH_TWO_PI)
            offset = depth * math.sin(phase)
            dsmp = (0.001 * (depth + offset)) * sr
            out[i] = (wa * s) + (wb * _fractional_delay(x, buf, i, dsmp))
        self.previous_phase = go_mod(prev + (ang * (float(len(buf)) * sr_inv)), MATH_TWO_PI)
        _push_history(buf, x)
        return out


class Phaser(Flanger):                    # effects/phaser.go:19-125: the flanger with a phase-dependent dry / wet split
    NUMERIC = {"depth": (0, 100, 100), "speed": (1, 100, 10), "phase": (-90, 90, 45)}

    def weights(self):
        fac = 0.5 * math.sin(MATH_DEGREE_TO_RADIANS * float(self.num["phase"]))
        return 1.0 - abs(fac), fac


class Tremolo(Unit):                      # effects/tremolo.go:15-65
    NUMERIC = {"frequency": (10, 100, 100), "phase": (0, 100, 50), "depth": (-60, 0, -10)}

    def __init__(self):
        super().__init__()
        self.attenuated = False
        self.in_state_since = 0

    def process(self, x, fs):
        period_f = float(fs) / (0.1 * float(self.num["frequency"]))
        period = u32(period_f)
        un = u32(period_f * (0.01 * float(self.num["phase"])))
        att_len = (period - un) & U32_MAX
        fac = decibels_to_factor(self.num["depth"])
        attenuated, since = self.attenuated, self.in_state_since
        out = [0.0] * len(x)
        for i, s in enumerate(x):
            r = s
            if attenuated and since >= att_len:
                attenuated, since = False, 0
            elif (not attenuated) and since >= un:
                attenuated, since = True, 0
            if attenuated:
                r *= fac
            out[i] = r
            since = (since + 1) & U32_MAX
        self.attenuated, self.in_state_since = attenuated, since
        return out


class RingModulator(Unit):                # effects/ringmodulator.go:18-47
    NUMERIC = {"frequency": (1, 100, 100)}

    def __init__(self):
        super().__init__()
        self.phase = 0.0

    def process(self, x, fs):
        phase = self.phase
        frac = (MATH_TWO_PI * float(self.num["frequency"])) / float(fs)
        out = [math.sin(go_mod(phase + (float(i) * frac), MATH_TWO_PI)) * s for i, s in enumerate(x)]
        self.phase = go_mod(phase + (float(len(x)) * frac), MATH_TWO_PI)
        return out


class Delay(Unit):                        # effects/delay.go:17-91
    NUMERIC = {"delay_time": (0, 1000, 200), "feedback": (-60, 0, -5), "level": (-30, 0, -5)}

    def __init__(self):
        super().__init__()
        self.buffer = []

    def process(self, x, fs):
        ds = int(math.floor(((0.001 * float(self.num["delay_time"])) * float(fs)) + 0.5))
        fb, lvl = decibels_to_factor(self.num["feedback"]), decibels_to_factor(self.num["level"])
        if len(self.buffer) != ds:
            self.buffer = [0.0] * ds
        buf = self.buffer
        out = [0.0] * len(x)
        for i, s in enumerate(x):
            out[i] = clamp1(lvl * (s + (fb * _delayed(x, buf, i - ds))))
        _push_history(buf, x)
        return out


class Reverb(Unit):                       # effects/reverb.go:41-116, 179-338
    NUMERIC = {"mix": (0, 100, 50)}

    def __init__(self):
        super().__init__()
        self.rate = 0
        self.allpasses = []               # [buffer, ptr, feedback]
        self.line = None                  # [buffer, indices, factors]

    @staticmethod
    def _go_round(v):                     # math.Round: half away from zero
        return math.floor(v + 0.5) if v >= 0 else -math.floor(-v + 0.5)

    def process(self, x, fs):
        n = len(x)
        wet = 0.01 * float(self.num["mix"])
        dry = 1.0 - wet
        sr = float(fs)
        if self.rate != fs:
            self.allpasses = [[[0.0] * int(self._go_round(dsec * sr)), 0, 0.7] for dsec in (0.04204, 0.01348, 0.00452)]
            idx = [u32(self._go_round(t * sr)) for t in (0.19196, 0.19996, 0.21596, 0.23204)]
            self.line = [[0.0] * max(idx), idx, [0.1855, 0.18325, 0.17875, 0.17425]]
            self.rate = fs
        # reverbDelayLine.process (reverb.go:65-116)
        lbuf, indices, factors = self.line
        bsize = len(lbuf)
        line_out = [0.0] * n
        for i in range(n):
            pre = 0.0
            for j, off in enumerate(indices):
                k = i - int(off)
                cur = 0.0
                if k >= 0:
                    cur = x[k]
                elif k >= -bsize:
                    cur = lbuf[bsize + k]
                pre += factors[j] * cur
            line_out[i] = pre
        _push_history(lbuf, x)
        front = list(line_out)
        for ap in self.allpasses:         # reverbAllpass.process (reverb.go:41-60)
            buf, ptr_w, fb = ap
            size = len(buf)
            back = [0.0] * n
            for i, s in enumerate(front):
                ptr_r = (ptr_w + 1) % size
                dl = buf[ptr_r]
                pre = s - (fb * dl)
                buf[ptr_w] = pre
                back[i] = (fb * pre) + dl
                ptr_w = ptr_r
            ap[1] = ptr_w
            front = back
        half_wet = 0.5 * wet
        return [clamp1((dry * x[i]) + (half_wet * (line_out[i] + front[i]))) for i in range(n)]


STRING_NONE = "- NONE -"


class PowerAmp(Unit):                     # effects/poweramp.go:25-216
    DISCRETE = {"filter_order": ([str(64 << k) for k in range(15)], 14)}

    def __init__(self, responses=None):
        self.DISCRETE = dict(PowerAmp.DISCRETE)
        self.NUMERIC = {}
        if responses is not None:         # PreparePowerAmp (poweramp.go:221-289)
            ext = [STRING_NONE] + responses.names()
            for i in range(1, NUM_FILTERS + 1):
                self.DISCRETE["filter_%d" % i] = (ext, 0)
                self.NUMERIC["level_%d" % i] = (-60, 0, 0)
        super().__init__()
        self.responses = responses
        self.rate = 0
        self.current = None

    def compile(self, rate):              # poweramp.go:25-127 -> Filter or None (error)
        irs = self.responses
        if irs is None:
            return None
        target = int(self.disc["filter_order"])
        filters = []
        for i in range(1, NUM_FILTERS + 1):
            name = self.disc["filter_%d" % i]
            flt = None
            if name != STRING_NONE:
                fac = decibels_to_factor(self.num["level_%d" % i])
                flt = irs.create_filter(name, rate)
                if flt is None:
                    return None
                if target > 0:
                    flt = flt.reduce(target)
                flt = flt.normalize()
                flt = flt.multiply(fac)
            filters.append(flt)
        comp = Filter([], rate, 0.0)      # filter.Empty
        for flt in filters:
            comp = comp.add(flt)
        return comp

    def _recompile(self):                 # SetDiscreteValue / SetNumericValue: poweramp.go:132-181
        flt = self.compile(self.rate)
        if flt is not None:
            self.current = flt

    def set_numeric(self, name, value):
        super().set_numeric(name, value)
        self._recompile()

    def set_discrete(self, name, value):
        super().set_discrete(name, value)
        self._recompile()

    def process(self, x, fs):             # poweramp.go:186-216
        if fs != self.rate:
            self.rate = fs
            flt = self.compile(fs)
            if flt is not None:
                self.current = flt
        if self.current is not None:
            return [float(v) for v in self.current.process(x)]
        return [0.0] * len(x)


# ---- effects.CreateUnit (effects/effects.go:443-516), unit type codes :21-43, names :543-565 ---------------------
UNIT_TYPES = ["signal_generator", "noise_gate", "bandpass", "auto_wah", "auto_yoy", "compressor", "octaver", "excess", "fuzz",
              "overdrive", "distortion", "tone_stack", "chorus", "flanger", "phaser", "tremolo", "ring_modulator", "delay",
              "reverb", "power_amp", "cabinet"]
_CLASSES = [SignalGenerator, NoiseGate, Bandpass, AutoWah, AutoYoy, Compressor, Octaver, Excess, Fuzz, Overdrive, Distortion,
            ToneStack, Chorus, Flanger, Phaser, Tremolo, RingModulator, Delay, Reverb, PowerAmp, Cabinet]


class Chain:                              # signal.chainStruct (signal/signal.go:41-47, 52-86, 361-414)
    def __init__(self, responses=None):
        self.responses = responses
        self.slots = []                   # [unit, bypass]

    def append_unit(self, unit_type):
        code = UNIT_TYPES.index(unit_type) if isinstance(unit_type, str) else unit_type
        cls = _CLASSES[code]
        unit = cls(self.responses) if cls is PowerAmp else cls()
        self.slots.append([unit, True])   # units start bypassed (signal.go:74)
        return len(self.slots) - 1

    def set_bypass(self, i, value):
        self.slots[i][1] = bool(value)

    def set_numeric_value(self, i, name, value):
        self.slots[i][0].set_numeric(name, value)

    def set_discrete_value(self, i, name, value):
        self.slots[i][0].set_discrete(name, value)

    def process(self, x, fs):
        buf = [float(v) for v in x]
        for unit, bypass in self.slots:
            if not bypass:
                buf = unit.process(buf, fs)
        return np.asarray(buf, dtype=np.float64)


# ---- spatializer (spatializer/spatializer.go) ------------------------------------------------------------
SPAT_EFFECTIVE_DISTANCE = _c(Fraction("0.215"))
SPAT_HALF_EFFECTIVE_DISTANCE = _c(Fraction(1, 2) * Fraction("0.215"))
SPAT_GROUP_DELAY = _c(Fraction("6.3e-4"))
SPAT_DELAY_PER_METRE = _c(Fraction("6.3e-4") / Fraction("0.215"))       # (GROUP_DELAY / EFFECTIVE_DISTANCE): one constant expression
SPAT_DEFAULT_RATE = 96000


class Spatializer:                        # spatializer.go:140-335, 418-469
    def __init__(self, channels):
        self.count = channels
        self.rate = SPAT_DEFAULT_RATE
        self.pos = [[0.0, 0.0, 1.0] for _ in range(channels)]     # azimuth, distance, level
        size = int(math.ceil(float(SPAT_DEFAULT_RATE) * SPAT_GROUP_DELAY))
        self.buffers = [[0.0] * size for _ in range(channels)]

    def set_sample_rate(self, rate):      # spatializer.go:418-431: re-makes the histories, does NOT store the rate
        size = int(math.ceil(float(rate) * SPAT_GROUP_DELAY))
        self.buffers = [[0.0] * size for _ in range(self.count)]

    def process(self, inputs, aux=None):
        n = len(inputs[0]) if len(inputs) else 0
        out_l, out_r = [0.0] * n, [0.0] * n
        sr = float(self.rate)
        for c, x in enumerate(inputs):
            az_deg, dist, level = self.pos[c]
            az = MATH_DEGREE_TO_RADIANS * az_deg
            buf = self.buffers[c]
            size = len(buf)
            xp, yp = dist * math.sin(az), dist * math.cos(az)
            xl, xr = abs(xp + SPAT_HALF_EFFECTIVE_DISTANCE), abs(xp - SPAT_HALF_EFFECTIVE_DISTANCE)
            yd = abs(yp)
            yd2 = yd * yd
            dl = math.sqrt((xl * xl) + yd2)
            with np.errstate(all="ignore"):
                pl = float(np.float64(1.0) / np.float64(dl))
            pl = 1.0 if pl > 1.0 else pl
            fl = level * pl
            dr = math.sqrt((xr * xr) + yd2)
            with np.errstate(all="ignore"):
                pr = float(np.float64(1.0) / np.float64(dr))
            pr = 1.0 if pr > 1.0 else pr
            fr = level * pr
            delay_time = SPAT_DELAY_PER_METRE * (dl - dr)
            dsmp = abs(delay_time) * sr
            early, late = math.floor(dsmp), math.ceil(dsmp)
            ie, il = int(early), int(late)
            if ie >= size:
                ie = size - 1
            if il >= size:
                il = size - 1
            for j in range(len(x)):
                cur = x[j]
                if delay_time == 0.0:
                    out_l[j] += fl * cur
                    out_r[j] += fr * cur
                else:
                    se = _delayed(x, buf, j - ie)
                    sl = _delayed(x, buf, j - il)
                    dsamp = ((1.0 - (dsmp - early)) * se) + ((1.0 - (late - dsmp)) * sl)
                    if delay_time > 0.0:
                        out_l[j] += fl * dsamp
                        out_r[j] += fr * x[j]
                    else:
                        out_l[j] += fl * x[j]
                        out_r[j] += fr * dsamp
        if aux is not None:
            for j, s in enumerate(aux):
                out_l[j] += s
                out_r[j] += s
        for c, x in enumerate(inputs):
            _push_history(self.buffers[c], [float(v) for v in x])
        return np.asarray(out_l), np.asarray(out_r)
