"""Pin the C oracle (oracle/) for the rows the reference holds no test vector for (SURVEY.md 8c ii / iii):

(ii)  tests/goref.py - a second restatement written independently from the Go sources - is diffed against the
      oracle at <= 1e-12 on every unit (defaults, parameter extremes, every discrete value, random draws), on
      CHAIN12, on the IR algebra (Reduce / Normalize / Multiply / Add, filter.Import resampling) and on the
      spatializer.  A misreading of a unit shared by the oracle, the host emulation and the CUDA stages (all written
      by one reader) would show here.
(iii) analytic spot checks of the oracle: closed forms that follow from the reference's formulas.

CPU only (-m "not gpu").
"""
import math
import zlib

import numpy as np
import pytest

import common
import goref
from common import orc

TOL = 1e-12                        # absolute, on signals of order 0.1 .. 1


def rel_ok(ref, got, tol=TOL, what=""):
    ref, got = np.asarray(ref, dtype=np.float64), np.asarray(got, dtype=np.float64)
    assert ref.shape == got.shape, what
    assert np.all(np.isfinite(got)), what
    err = np.max(np.abs(ref - got)) if ref.size else 0.0
    assert err <= tol * max(1.0, np.max(np.abs(ref)) if ref.size else 1.0), "%s: max |diff| %.3e (peak %.3f)" % (what, err, np.max(np.abs(ref)))
    return err


def test_signal(fs, n_total, seed, amp=0.7):
    """two partials + noise with a loud/quiet envelope: crosses gate / compressor thresholds and the clamps"""
    rng = np.random.default_rng(seed)
    t = np.arange(n_total) / fs
    x = 0.5 * np.sin(2 * np.pi * 196.0 * t) + 0.25 * np.sin(2 * np.pi * 1234.5 * t + 0.3) + 0.05 * rng.uniform(-1, 1, n_total)
    env = 0.05 + 0.95 * (np.sin(2 * np.pi * 3.0 * t) > -0.3)
    return amp * x * env


test_signal.__test__ = False


def build_pair(spec, banks=None, ir_name=None):
    ob, gb = banks if banks else (None, None)
    a = common.build_chain(orc.Chain(ob), spec, ir_name)
    b = common.build_chain(goref.Chain(gb), spec, ir_name)
    return a, b


def stream_pair(a, b, x, n, fs, what, tol=TOL):
    worst = 0.0
    any_signal = False
    for k in range(0, len(x), n):
        ya, yb = a.process(x[k:k + n], fs), b.process(x[k:k + n], fs)
        worst = max(worst, rel_ok(yb, ya, tol, "%s block %d" % (what, k // n)))
        any_signal = any_signal or np.max(np.abs(yb)) > 0
    return worst, any_signal


# ---------------------------------------------------------------------------------------------------------------
# (ii) every unit: parameter tables and Process bodies
# ---------------------------------------------------------------------------------------------------------------
UNITS = [u for u in goref.UNIT_TYPES if u != "power_amp"]
PANIC_RATES = {"flanger": (22050, 44100, 88200), "phaser": (22050, 44100, 88200)}    # see test_flanger_reference_panics


def unit_variants(name):
    """[(numeric, discrete)]: defaults, all-min, all-max, each discrete value, 3 seeded random draws"""
    cls = goref._CLASSES[goref.UNIT_TYPES.index(name)]
    num, disc = cls.NUMERIC, cls.DISCRETE
    out = [({}, {})]
    if num:
        out.append(({k: v[0] for k, v in num.items()}, {}))
        out.append(({k: v[1] for k, v in num.items()}, {}))
    for k, (vals, _d) in disc.items():
        for v in vals:
            out.append(({}, {k: v}))
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    for _ in range(3):
        out.append(({k: int(rng.integers(v[0], v[1] + 1)) for k, v in num.items()},
                    {k: vals[int(rng.integers(0, len(vals)))] for k, (vals, _d) in disc.items()}))
    return out


def test_parameter_tables_match_oracle():
    """names, ranges and defaults of every unit: goref's tables were extracted from the Go Parameter{} literals, the
    oracle's come from include/gdg_unit_params.def"""
    for name in UNITS:
        cls = goref._CLASSES[goref.UNIT_TYPES.index(name)]
        c = orc.Chain()
        i = c.append_unit(name)
        for k, (mn, mx, dflt) in cls.NUMERIC.items():
            assert c.get_numeric_value(i, k) == dflt, (name, k)
            c.set_numeric_value(i, k, mn)
            c.set_numeric_value(i, k, mx)
            for bad in (mn - 1, mx + 1):
                with pytest.raises(orc.OracleError):
                    c.set_numeric_value(i, k, bad)
        for k, (vals, d) in cls.DISCRETE.items():
            assert c.get_discrete_value(i, k) == vals[d], (name, k)
            for v in vals:
                c.set_discrete_value(i, k, v)


@pytest.mark.parametrize("name", UNITS)
@pytest.mark.parametrize("fs,n,blocks", [(32000, 512, 10), (192000, 256, 3), (22050, 64, 24)])
def test_unit_crosscheck(name, fs, n, blocks):
    if fs in PANIC_RATES.get(name, ()):
        fs = 48000
    x = test_signal(fs, n * blocks, seed=zlib.crc32(("%s%d" % (name, fs)).encode()) % 1000)
    for vi, (num, disc) in enumerate(unit_variants(name)):
        if name in ("excess", "fuzz", "overdrive", "distortion") and fs == 192000 and vi > 6:
            continue                                           # the sequential Python up-sampler is slow: fewer variants at 4 x 192 kHz
        a, b = build_pair([(name, num, disc)])
        _w, _ = stream_pair(a, b, x, n, fs, "%s %s %s @%d" % (name, num, disc, fs))


def test_reverb_taps_are_reached():
    """enough blocks at a low rate that the 192-232 ms delay-line taps and all three all-passes contribute"""
    fs, n = 22050, 1024
    x = test_signal(fs, 10 * n, 5)
    a, b = build_pair([("reverb", {"mix": 70}, {})])
    worst, nz = stream_pair(a, b, x, n, fs, "reverb long")
    assert nz
    dry = common.build_chain(orc.Chain(), [("reverb", {"mix": 0}, {})])
    y = np.concatenate([dry.process(x[k:k + n], fs) for k in range(0, len(x), n)])
    assert np.array_equal(y, np.clip(x, -1, 1))                # mix 0: dry * 1.0 + 0.0 * wet, clamp only


def test_delay_and_chorus_state_carry_short_blocks():
    """blocks much shorter than the history: the history shift path (boundary >= 0) of every delay-type unit"""
    fs, n = 48000, 32
    x = test_signal(fs, n * 40, 9)
    for spec in ([("delay", {"delay_time": 3, "feedback": -3, "level": 0}, {})], [("chorus", {"depth": 100, "speed": 100}, {})],
                 [("auto_yoy", {"depth": 100}, {"follow": "envelope"})], [("flanger", {"depth": 100, "speed": 100}, {})],
                 [("phaser", {"depth": 100, "speed": 100, "phase": -90}, {})]):
        a, b = build_pair(spec)
        stream_pair(a, b, x, n, fs, str(spec))


def test_oversampler_block_length_change():
    """the up-sampler re-makes (zeroes) its 8-sample history when the block length changes (oversampling.go:74-84)"""
    fs = 44100
    x = test_signal(fs, 2048, 3)
    a, b = build_pair([("distortion", {"gain": 12}, {"oversampling": "4"})])
    k = 0
    for n in (256, 256, 128, 512, 64, 64, 256):
        rel_ok(b.process(x[k:k + n], fs), a.process(x[k:k + n], fs), TOL, "os block %d" % n)
        k += n


def test_chain12_crosscheck():
    """the canonical 12-stage chain with a synthetic 2048-tap IR reduced to 512 taps"""
    fs, n, blocks = 32000, 512, 8
    ir = common.synthetic_ir(2048)
    ob, gb = orc.IRBank(), goref.ImpulseResponses()
    ob.add_raw("syn", -3, ir, fs)
    gb.add_raw("syn", -3, ir, fs)
    spec = [(u, dict(nn), dict(dd)) for u, nn, dd in common.CHAIN12]
    for s in spec:
        if s[0] == "power_amp":
            s[2]["filter_order"] = "512"
    a, b = build_pair(spec, (ob, gb), "syn")
    x = test_signal(fs, n * blocks, 77, amp=0.9)
    worst, nz = stream_pair(a, b, x, n, fs, "CHAIN12")
    assert nz


def test_all_units_one_chain():
    fs, n = 48000, 256
    spec = [(u, {}, {}) for u in UNITS]
    a, b = build_pair(spec)
    x = test_signal(fs, n * 6, 21)
    stream_pair(a, b, x, n, fs, "all units", tol=1e-11)


def test_bypass_and_order():
    """signal.Chain.Process: bypassed slots are skipped, slot order is processing order (signal.go:361-414)"""
    fs, n = 44100, 128
    x = test_signal(fs, n * 3, 2)
    spec = [("distortion", {"gain": 20}, {}), ("tone_stack", {}, {}), ("tremolo", {}, {})]
    a, b = build_pair(spec)
    a.set_bypass(1, True)
    b.set_bypass(1, True)
    stream_pair(a, b, x, n, fs, "bypass")
    empty_a, empty_b = orc.Chain(), goref.Chain()
    empty_a.append_unit("fuzz")
    empty_b.append_unit("fuzz")                                # appended units start bypassed (signal.go:74)
    assert np.array_equal(empty_a.process(x[:n], fs), x[:n]) and np.array_equal(empty_b.process(x[:n], fs), x[:n])


# ---------------------------------------------------------------------------------------------------------------
# (ii) IR algebra and the power amp
# ---------------------------------------------------------------------------------------------------------------
def test_filter_algebra_crosscheck():
    rng = np.random.default_rng(11)
    h = rng.uniform(-1, 1, 1500) * np.exp(-np.arange(1500) / 300.0)
    fo, fg = orc.Filter(h, 48000), goref.Filter(h, 48000, 0.0)
    fo.set_compensation(1.4125375446227544)
    fg.comp = 1.4125375446227544
    for order in (64, 256, 1024, 2048):
        ro, rg = fo.reduce(order), fg.reduce(order)
        rel_ok(rg.coeffs, ro.coefficients(), 1e-12, "Reduce(%d)" % order)
        no, ng = ro.normalize(), rg.normalize()
        rel_ok(ng.coeffs, no.coefficients(), 1e-12, "Normalize after Reduce(%d)" % order)
        assert abs(math.sqrt(sum(c * c for c in ng.coeffs)) - 1.4125375446227544) < 1e-12      # ||h||_2 = compensation
        mo, mg = no.multiply(0.25), ng.multiply(0.25)
        rel_ok(mg.coeffs, mo.coefficients(), 1e-13, "Multiply")
        assert mg.comp == 0.0                                   # filter.go:299-304: Multiply drops the compensation
    x = test_signal(48000, 2048, 4)
    yo = np.concatenate([fo.process(x[k:k + 512]) for k in range(0, 2048, 512)])
    yg = np.concatenate([fg.process(x[k:k + 512]) for k in range(0, 2048, 512)])
    rel_ok(yg, yo, 1e-12, "filter.Process, L=1500 > N=512")
    assert np.max(np.abs(yo)) == 1.0                            # the clamp is active in this case
    direct = np.clip(np.convolve(x, h)[:2048], -1, 1)
    # clamp only on the way out (filter.go:473-485): overlap-add of unclamped tails = linear convolution, then clamp
    rel_ok(direct, yo, 1e-11, "filter.Process vs direct convolution")


def test_filter_import_resampling_crosscheck():
    rng = np.random.default_rng(5)
    src = rng.uniform(-1, 1, 300) * np.exp(-np.arange(300) / 60.0)
    ob, gb = orc.IRBank(), goref.ImpulseResponses()
    ob.add("ir", -6, src, 96000)
    gb.import_one("ir", -6, src, 96000)
    assert ob.names() == gb.names() == ["ir"]
    for rate in goref.SAMPLE_RATES:
        spec = [("power_amp", {"level_1": -3}, {"filter_order": "1048576", "filter_1": "ir"})]
        a, b = build_pair(spec, (ob, gb))
        x = test_signal(rate, 1024, rate % 97)
        stream_pair(a, b, x, 256, rate, "power_amp import @%d" % rate)
        flt = gb.create_filter("ir", rate)
        expect = int(math.floor(300 * rate / 96000.0)) - (1 if (300 * rate) % 96000 == 0 else 0)
        assert len(flt.coeffs) == expect, (rate, len(flt.coeffs))


def test_poweramp_mix_and_recompile_crosscheck():
    fs, n = 44100, 256
    rng = np.random.default_rng(8)
    ob, gb = orc.IRBank(), goref.ImpulseResponses()
    for name, L in (("A", 700), ("B", 1900), ("C", 90)):
        h = rng.uniform(-1, 1, L) * np.exp(-np.arange(L) * 5.0 / L)
        ob.add_raw(name, int(L % 7) - 3, h, fs)
        gb.add_raw(name, int(L % 7) - 3, h, fs)
    spec = [("power_amp", {"level_1": -2, "level_2": -9, "level_3": 0}, {"filter_order": "1024", "filter_1": "A", "filter_2": "B", "filter_3": "C"})]
    a, b = build_pair(spec, (ob, gb))
    x = test_signal(fs, n * 12, 6, amp=0.4)
    stream_pair(a, b, x[:4 * n], n, fs, "power_amp mix")
    for ch in (a, b):                                          # a parameter change recompiles: new filter object, fresh tail (poweramp.go:132-181)
        ch.set_numeric_value(0, "level_2", -20)
    stream_pair(a, b, x[4 * n:8 * n], n, fs, "power_amp after level change")
    for ch in (a, b):
        ch.set_discrete_value(0, "filter_1", "- NONE -")
        ch.set_discrete_value(0, "filter_2", "- NONE -")
        ch.set_discrete_value(0, "filter_3", "- NONE -")
    ya, yb = a.process(x[8 * n:9 * n], fs), b.process(x[8 * n:9 * n], fs)
    assert not ya.any() and not yb.any()                       # empty composite filter -> zeros (filter.go:366-367)


def test_poweramp_without_bank_or_rate():
    fs, n = 48000, 128
    x = test_signal(fs, n, 1)
    a, b = build_pair([("power_amp", {}, {})])
    assert not a.process(x, fs).any() and not b.process(x, fs).any()      # no filter compiled: zeros (poweramp.go:210-214)


# ---------------------------------------------------------------------------------------------------------------
# (ii) spatializer
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rate", [96000, 192000, 44100])
def test_spatializer_crosscheck(rate):
    C, n = 7, 96
    so, sg = orc.Spatializer(C), goref.Spatializer(C)
    if rate != 96000:
        so.set_sample_rate(rate)
        sg.set_sample_rate(rate)
    geo = [(-90.0, 0.3, 1.0), (-40.0, 1.0, 0.8), (0.0, 2.0, 1.0), (0.0, 0.0, 0.5), (25.0, 0.05, 1.0), (90.0, 10.0, 0.25), (180.0, 4.0, 1.0)]
    for c, (az, dist, lvl) in enumerate(geo):
        so.set_azimuth(c, az), so.set_distance(c, dist), so.set_level(c, lvl)
        sg.pos[c] = [az, dist, lvl]
    rng = np.random.default_rng(3)
    for k in range(4):
        x = rng.uniform(-1, 1, (C, n))
        aux = rng.uniform(-0.1, 0.1, n) if k % 2 else None
        yo = so.process(x, aux)
        gl, gr = sg.process([list(r) for r in x], list(aux) if aux is not None else None)
        rel_ok(gl, yo[0], 1e-13, "spat L block %d" % k)
        rel_ok(gr, yo[1], 1e-13, "spat R block %d" % k)


# ---------------------------------------------------------------------------------------------------------------
# the reference's out-of-range read (flanger / phaser at 22.05, 44.1 and 88.2 kHz)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("unit", ["flanger", "phaser"])
@pytest.mark.parametrize("fs,n", [(22050, 32), (44100, 64), (88200, 128)])
def test_flanger_reference_panics_oracle_clamps(unit, fs, n):
    """At depth 100 the delay reaches 0.002 fs (44.1 / 88.2 / 176.4 samples) but the history holds round(0.002 fs) = 44 /
    88 / 176, so ceil(delay) = history + 1 and the late tap of sample 0 asks for buffer[-1]: the reference panics
    (flanger.go:80-84).  The literal restatement raises; the oracle (and the product) deliberately read the OLDEST history
    sample instead and never touch memory outside the history."""
    speed = 37
    c = goref.Chain()
    c.append_unit(unit), c.set_numeric_value(0, "speed", speed), c.set_bypass(0, False)
    x = test_signal(fs, n * 600, 12)
    with pytest.raises(goref._GoIndexPanic):
        for k in range(0, len(x), n):
            c.process(x[k:k + n], fs)

    class Clamped(goref.Chain):                                # the literal restatement with the oracle's documented clamp
        pass
    orig = goref._delayed
    goref._delayed = lambda xx, buf, idx: orig(xx, buf, max(idx, -len(buf)))
    try:
        a = common.build_chain(orc.Chain(), [(unit, {"speed": speed}, {})])
        b = common.build_chain(Clamped(), [(unit, {"speed": speed}, {})])
        stream_pair(a, b, x, n, fs, "%s clamped @%d" % (unit, fs))
    finally:
        goref._delayed = orig


# ---------------------------------------------------------------------------------------------------------------
# (iii) analytic spot checks of the oracle
# ---------------------------------------------------------------------------------------------------------------
def run(spec, x, fs, n=None):
    c = common.build_chain(orc.Chain(), spec)
    n = n or len(x)
    return np.concatenate([c.process(x[k:k + n], fs) for k in range(0, len(x), n)])


def test_analytic_distortion_is_exact_clamp():
    x = np.linspace(-3, 3, 1001)
    assert np.array_equal(run([("distortion", {"boost": 0, "gain": 0, "level": 0}, {})], x, 48000), np.clip(x, -1, 1))
    g = 10 ** (0.05 * 12)
    y = run([("distortion", {"boost": 6, "gain": 6, "level": -6}, {})], x, 48000)
    assert np.allclose(y, 10 ** (0.05 * -6) * np.clip(g * x, -1, 1), rtol=0, atol=2e-16)


def test_analytic_overdrive_ecc83_is_tanh():
    x = np.linspace(-1, 1, 801)
    g = 10 ** (0.05 * 20)
    y = run([("overdrive", {"gain": 20, "boost": 0, "drive": 100, "level": 0}, {"valve": "ECC83 (12AX7)"})], x, 48000)
    assert np.max(np.abs(y - np.tanh(0.5 * g * x))) < 5e-16      # 2/(1+e^-a) - 1 = tanh(a/2)
    y82 = run([("overdrive", {"gain": 0, "boost": 0, "drive": 50, "level": 0}, {"valve": "ECC82 (12AU7)"})], x, 48000)
    assert np.max(np.abs(y82 - (0.5 * (2 / np.pi) * np.arctan(0.25 * np.pi * x) + 0.5 * x))) < 5e-16


def test_analytic_excess_is_triangle_fold():
    x = np.linspace(-1, 1, 1999)
    g = 10 ** (0.05 * 30)
    y = run([("excess", {"gain": 30, "level": 0}, {})], x, 48000)
    tri = (2 / np.pi) * np.arcsin(np.sin(0.5 * np.pi * g * x))   # triangular fold of g x into [-1, 1]
    assert np.max(np.abs(y - tri)) < 1e-12
    small = np.linspace(-0.03, 0.03, 101)
    assert np.array_equal(run([("excess", {"gain": 30, "level": 0}, {})], small, 48000), g * small)     # |g x| <= 1: untouched


def test_analytic_linear_units():
    fs, n = 48000, 512
    x = test_signal(fs, n * 4, 14, amp=0.2)
    for unit in ("cabinet", "tone_stack", "bandpass"):
        y1 = run([(unit, {}, {})], x, fs, n)
        y2 = run([(unit, {}, {})], 0.5 * x, fs, n)
        assert np.max(np.abs(y1 - 2 * y2)) < 1e-15, unit       # linear while unclamped: scaling by 0.5 is exact
        dc = run([(unit, {}, {})], np.full(fs // 2, 0.3), fs, 1024)
        assert abs(dc[-1]) < 1e-3, unit                         # every path has a high-pass: DC is blocked
    # one-pole low-pass: cabinet's last section at 6 kHz etc. -> a 100 Hz sine passes the LP side almost untouched
    t = np.arange(fs) / fs
    y = run([("tone_stack", {"low": 0, "middle": -30, "presence": -30, "high": -30}, {})], 0.5 * np.sin(2 * np.pi * 100 * t), fs, 1024)
    band = np.abs(y[fs // 2:]).max()
    # band 1 = HP 20 Hz -> LP 300 Hz at 100 Hz: |H| = (f/f1)/sqrt(1+(f/f1)^2) / sqrt(1+(f/f2)^2)
    assert abs(band - 0.5 * (5 / math.sqrt(26)) / math.sqrt(1 + 1 / 9.0)) < 0.03


def test_analytic_tremolo_ringmod_delay():
    fs = 48000
    x = np.ones(fs // 2)
    y = run([("tremolo", {"frequency": 100, "phase": 50, "depth": -20}, {})], x, fs, 1024)
    period = int(fs / 10.0)
    un = int((fs / 10.0) * 0.5)
    assert np.allclose(y[:un], 1.0) and np.allclose(y[un:period], 0.1)          # unattenuated first, then depth for the rest of the period
    assert np.array_equal(y[period:2 * period], y[:period])
    r = run([("ring_modulator", {"frequency": 100}, {})], x, fs, 1024)
    assert np.max(np.abs(r - np.sin(2 * np.pi * 100 * np.arange(len(x)) / fs))) < 1e-9
    imp = np.zeros(4096)
    imp[0] = 0.5
    d = run([("delay", {"delay_time": 10, "feedback": -6, "level": 0}, {})], imp, fs, 256)
    assert d[0] == 0.5 and abs(d[480] - 0.5 * 10 ** (-0.3)) < 1e-16 and np.count_nonzero(d) == 2       # feed-forward echo, one tap


def test_analytic_chorus_depth_zero_is_comb():
    fs, n = 48000, 256
    x = test_signal(fs, n * 12, 4)
    y = run([("chorus", {"depth": 0, "speed": 1}, {})], x, fs, n)
    d = int(0.04 * fs)                                          # 40 ms, integral: both interpolation weights are 1 (chorus.go:104-105)
    ref = 0.5 * x.copy()
    ref[d:] += 0.5 * (5 * 0.2 * 2.0 * x[:-d])
    assert np.max(np.abs(y - ref)) < 1e-15


def test_analytic_compressor_gain_law():
    fs = 48000
    x = np.full(fs, 0.01)
    y = run([("compressor", {"target_level": -20, "gain_limit": 30}, {"follow": "level"})], x, fs, 1024)
    assert abs(y[-1] - 0.1) < 1e-6                              # envelope -> 0.01, gain = 0.1 / 0.01 = 10 (< limit 31.6)
    y2 = run([("compressor", {"target_level": 0, "gain_limit": 6}, {"follow": "envelope"})], x, fs, 1024)
    assert abs(y2[-1] - 0.01 * 10 ** 0.3) < 1e-12               # limited


def test_analytic_noise_gate():
    fs = 48000
    x = np.concatenate([np.full(100, 0.5), np.full(5000, 0.001)])
    y = run([("noise_gate", {"threshold_open": -20, "threshold_close": -40, "hold_time": 50}, {})], x, fs, 100)
    hold = int(math.floor(0.05 * fs + 0.5))
    assert np.array_equal(y[:100], x[:100])
    assert np.array_equal(y[100:100 + hold - 1], x[100:100 + hold - 1]) and not y[100 + hold:].any()
