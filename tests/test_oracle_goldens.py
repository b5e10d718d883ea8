"""Pin the CPU oracle against every golden vector the reference's own tests hold
for this path.  Tolerances are the reference's own.

  fft/fft_test.go            1e-8   (fft_test.go:33)
  oversampling_test.go       1e-7   (oversampling_test.go:33)
  resample_test.go           1e-7
  random_test.go             1e-8
  circular_test.go           exact
  tuner_test.go              note exact, |cents| <= 5, finite frequency
"""
import glob
import os

import numpy as np
import pytest

import pyoracle as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def close(a, b, tol):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    assert np.max(np.abs(a - b)) <= tol, (a, b)


# ---- fft/fft_test.go -------------------------------------------------------
def test_next_power_of_two(goldens):
    g = goldens["fft/fft_test.go"]["TestNextPowerOfTwo"]
    for v, p, e in zip(g["in"], g["powers"], g["exponents"]):
        assert orc.next_power_of_two(v) == (p, e)


def test_real_fft(goldens):
    g = goldens["fft/fft_test.go"]["TestRealFFT"]
    assert len(g["in"]) == 7
    for x, re, im in zip(g["in"], g["outRealExpected"], g["outImagExpected"]):
        X = orc.real_fourier(x)
        close(X.real, re, 1e-8)
        close(X.imag, im, 1e-8)
        close(orc.real_inverse_fourier(X), x, 1e-8)


def test_complex_fft_both_modes(goldens):
    g = goldens["fft/fft_test.go"]["TestComplexFFT"]
    for mode in (orc.MODE_STANDARD, orc.MODE_INPLACE):
        for xr, xi, re, im in zip(g["inReal"], g["inImag"], g["outRealExpected"], g["outImagExpected"]):
            x = np.array(xr) + 1j * np.array(xi)
            X = orc.fourier(x, orc.SCALING_DEFAULT, mode)
            close(X.real, re, 1e-8)
            close(X.imag, im, 1e-8)
            y = orc.inverse_fourier(X, orc.SCALING_DEFAULT, mode)
            close(y.real, xr, 1e-8)
            close(y.imag, xi, 1e-8)


def test_orthonormal_scaling(goldens):
    g = goldens["fft/fft_test.go"]["TestOrthonormalScaling"]
    X = orc.real_fourier(g["in"], orc.SCALING_ORTHONORMAL)
    close(X.real, g["expectedReal"], 1e-8)
    close(X.imag, g["expectedImag"], 1e-8)
    close(orc.real_inverse_fourier(X, orc.SCALING_ORTHONORMAL), g["in"], 1e-8)


def test_single_element_and_failure_cases(goldens):
    g = goldens["fft/fft_test.go"]["TestSingleElementFFT"]
    X = orc.real_fourier(g["inReal"])
    assert X[0] == complex(g["inReal"][0], 0.0)
    assert orc.real_inverse_fourier(X)[0] == g["inReal"][0]
    with pytest.raises(orc.OracleError):
        orc.real_fourier([1.0, 2.0, 3.0])
    with pytest.raises(orc.OracleError):
        orc.real_inverse_fourier([1.0, 2.0, 3.0])


def test_shift(goldens):
    g = goldens["fft/fft_test.go"]["TestShift"]
    c = lambda v: np.array([complex(a, b) for a, b in v])
    assert np.array_equal(orc.shift(c(g["inEven"])), c(g["outEven"]))
    assert np.array_equal(orc.shift(c(g["outEven"]), inverse=True), c(g["inEven"]))
    assert np.array_equal(orc.shift(c(g["inOdd"])), c(g["outOdd"]))
    assert np.array_equal(orc.shift(c(g["outOdd"]), inverse=True), c(g["inOdd"]))


# ---- oversampling/oversampling_test.go --------------------------------------
@pytest.mark.parametrize("factor,fn", [(2, "TestTwoTimesOversampling"), (4, "TestFourTimesOversampling")])
def test_oversampler_decimator_sequences(goldens, factor, fn):
    g = goldens["oversampling/oversampling_test.go"][fn]
    osd = orc.OversamplerDecimator(factor)
    assert len(g["in"]) == 4
    for x, up, down in zip(g["in"], g["oversampledExpected"], g["decimatedExpected"]):
        u = osd.oversample(x)
        close(u, up, 1e-7)
        close(osd.decimate(u), down, 1e-7)


# ---- resample/resample_test.go ------------------------------------------------
def test_resample_time(goldens):
    g = goldens["resample/resample_test.go"]["TestTimeSeries"]
    for x, up, down in zip(g["in"], g["outExpectedUp"], g["outExpectedDown"]):
        close(orc.resample_time(x, 96000, 192000), up, 1e-7)
        close(orc.resample_time(x, 96000, 44100), down, 1e-7)


def test_resample_frequency(goldens):
    g = goldens["resample/resample_test.go"]["TestFrequencySeries"]
    for x, re, im in zip(g["in"], g["outExpectedReal"], g["outExpectedImag"]):
        r = orc.resample_frequency([complex(a, b) for a, b in x], 4)
        close(r.real, re, 1e-7)
        close(r.imag, im, 1e-7)


def test_resample_oversample(goldens):
    g = goldens["resample/resample_test.go"]["TestOversample"]
    for x, e in zip(g["in"], g["outExpected"]):
        close(orc.resample_oversample(x, len(e), 2), e, 1e-7)


# ---- random/random_test.go -------------------------------------------------------
def test_prng(goldens):
    g = goldens["random/random_test.go"]["TestRNG"]
    for seed, exp in zip(g["seeds"], g["expectedOutputs"]):
        r = orc.PRNG(seed)
        close(r.floats(8), exp, 1e-8)
        more = r.floats(10000)
        assert more.min() >= 0.0 and more.max() <= 1.0


# ---- circular/circular_test.go ------------------------------------------------------
def test_circular(goldens):
    g = goldens["circular/circular_test.go"]["TestBuffer"]
    i, e = g["in"], g["expected"]
    b = orc.Circular(5)
    outs = []
    b.enqueue(*i[0]); outs.append(b.retrieve())
    b.enqueue(*i[1]); outs.append(b.retrieve())
    b.enqueue(*i[2]); outs.append(b.retrieve())
    b.enqueue(*i[3]); b.enqueue(*i[4]); outs.append(b.retrieve())
    b.enqueue(*i[5]); outs.append(b.retrieve())
    b.enqueue(*i[6]); outs.append(b.retrieve())
    for k in (7, 8, 9):
        b.enqueue(*i[k])
    b.enqueue(i[10][0]); b.enqueue(i[11][0]); outs.append(b.retrieve())
    for o, x in zip(outs, e):
        assert np.array_equal(o, x)
    with pytest.raises(orc.OracleError):
        b.retrieve(4)
    assert b.length() == 5


# ---- tuner/tuner_test.go ----------------------------------------------------------------
@pytest.mark.parametrize("note", ["D2", "A2", "D3", "G3", "H3", "E4"])
def test_tuner_notes(note):
    z = np.load(os.path.join(GOLD, "tuner_tail_%s.npz" % note))
    t = orc.Tuner()
    t.process(z["samples"], int(z["sample_rate"]))
    r = t.analyze()
    assert r["note"] == note
    assert -5 <= r["cents"] <= 5
    assert np.isfinite(r["frequency"])


def test_tuner_sequence_like_reference_test():
    """tuner_test.go feeds all six recordings through ONE tuner in order."""
    t = orc.Tuner()
    for note in ["D2", "A2", "D3", "G3", "H3", "E4"]:
        z = np.load(os.path.join(GOLD, "tuner_tail_%s.npz" % note))
        t.process(z["samples"], int(z["sample_rate"]))
        assert t.analyze()["note"] == note


# ---- PCM16 codec (wave/wave.go:347-428) ----------------------------------------------------
def test_pcm16_roundtrip_bit_exact():
    pcm = np.arange(-32768, 32768, dtype=np.int16)
    x = orc.pcm16_decode(pcm)
    assert x.min() >= -1.0001 and x.max() <= 1.0
    back = orc.pcm16_encode(x)
    # int32(32767.5 * (2/65535) * k) truncates toward zero: exact for k >= 0,
    # and for k < 0 as well because 32767.5 * 2 / 65535 == 1 up to rounding
    assert np.max(np.abs(back.astype(np.int32) - pcm.astype(np.int32))) <= 1
    assert orc.pcm16_encode([2.0, -2.0, 0.0])[0] == 32767
    assert orc.pcm16_encode([2.0, -2.0, 0.0])[1] == -32767


# ---- wave/wave_test.go: all six sample formats, byte-exact export and import within the test's tolerance -------
@pytest.mark.parametrize("fmt", ["PCM8", "PCM16", "PCM24", "PCM32", "IEEE32", "IEEE64"])
def test_wave_codecs(fmt):
    import json
    g = json.load(open(os.path.join(GOLD, "wave_goldens.json")))[fmt]
    enc = orc.wave_encode(fmt, g["export_samples"])
    assert list(enc) == g["export_data"]                      # TestExport<fmt>Mono: byte-identical payload
    dec = orc.wave_decode(fmt, bytes(g["import_data"]))
    close(dec, g["import_samples"], g["import_tolerance"])    # TestImport<fmt>Mono
    # the PCM16 helpers agree with the generic codec
    if fmt == "PCM16":
        pcm = np.frombuffer(bytes(g["import_data"]), dtype="<i2")
        assert np.array_equal(orc.pcm16_decode(pcm), dec)


def test_level_meter_matches_reference_test():
    """level/level_test.go TestMeters: -3 / 0 dB and -9 / -6 dB for a 1 Hz sine of amplitude 1 and 0.5, -200 once disabled."""
    import json
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "level_goldens.json")))
    fs, n = g["sample_rate"], g["samples"]
    sig = np.sin((2.0 * np.pi) * (np.arange(n, dtype=np.float64) / float(fs)))
    m = orc.Meter(2)
    m.process([sig, 0.5 * sig], fs)                           # created disabled (level.go:383): nothing happens
    assert [m.analyze(c) for c in range(2)] == [(-200, -200)] * 2
    m.set_enabled(True)
    m.process([a * sig for a in g["amplitudes"]], fs)
    assert [m.analyze(c)[0] for c in range(2)] == g["enabled"]["level"]
    assert [m.analyze(c)[1] for c in range(2)] == g["enabled"]["peak"]
    m.set_enabled(False)
    assert [m.analyze(c)[0] for c in range(2)] == g["after_disable"]["level"]
    assert [m.analyze(c)[1] for c in range(2)] == g["after_disable"]["peak"]
    with pytest.raises(orc.OracleError):
        m.process([sig], fs)                                  # level.go:311-312: buffer count mismatch
