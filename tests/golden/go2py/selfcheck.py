"""The translated packages against the reference's OWN unit-test vectors (tests/golden/reference_goldens.json, extracted from
fft_test.go, oversampling_test.go, resample_test.go, random_test.go, circular_test.go) at the reference's own tolerances:
this is what says the mechanical translation preserves the semantics of the Go sources."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def fl(x, kind='f'):
    from goruntime import G
    return G.lit([float(v) for v in x] if kind == 'f' else [complex(v) for v in x], kind)


def close(a, b, tol):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.max(np.abs(a - b)) <= tol, float(np.max(np.abs(a - b)))


def run(loader, verbose=False):
    from goruntime import G
    gold = json.load(open(os.path.join(HERE, "..", "reference_goldens.json")))
    n = 0
    fft = loader.load("fft")
    g = gold["fft/fft_test.go"]
    t = g["TestNextPowerOfTwo"]
    for v, p, e in zip(t["in"], t["powers"], t["exponents"]):
        assert tuple(fft.NextPowerOfTwo(v)) == (p, e); n += 1
    ft = fft.CreateFourierTransform()
    t = g["TestRealFFT"]
    for x, re, im in zip(t["in"], t["outRealExpected"], t["outImagExpected"]):
        out = G.make_slice(0j, len(x), None, 'c')
        assert ft.RealFourier(fl(x), out, fft.SCALING_DEFAULT) is None
        X = np.array(out.tolist())
        close(X.real, re, 1e-8); close(X.imag, im, 1e-8)
        back = G.make_slice(0.0, len(x), None, 'f')
        assert ft.RealInverseFourier(out, back, fft.SCALING_DEFAULT) is None
        close(back.tolist(), x, 1e-8); n += 1
    t = g["TestComplexFFT"]
    for mode in (fft.MODE_STANDARD, fft.MODE_INPLACE):
        for xr, xi, re, im in zip(t["inReal"], t["inImag"], t["outRealExpected"], t["outImagExpected"]):
            x = fl(np.array(xr) + 1j * np.array(xi), 'c')
            X = ft.Fourier(x, fft.SCALING_DEFAULT, mode)
            Xa = np.array(X.tolist())
            close(Xa.real, re, 1e-8); close(Xa.imag, im, 1e-8)
            y = np.array(ft.InverseFourier(X, fft.SCALING_DEFAULT, mode).tolist())
            close(y.real, xr, 1e-8); close(y.imag, xi, 1e-8); n += 1
    t = g["TestOrthonormalScaling"]
    out = G.make_slice(0j, len(t["in"]), None, 'c')
    assert ft.RealFourier(fl(t["in"]), out, fft.SCALING_ORTHONORMAL) is None
    X = np.array(out.tolist())
    close(X.real, t["expectedReal"], 1e-8); close(X.imag, t["expectedImag"], 1e-8); n += 1
    t = g["TestShift"]
    c = lambda v: [complex(a, b) for a, b in v]
    for a, b in (("inEven", "outEven"), ("inOdd", "outOdd")):
        v = fl(c(t[a]), 'c')
        fft.Shift(v, False)
        assert v.tolist() == c(t[b])
        fft.Shift(v, True)
        assert v.tolist() == c(t[a]); n += 1
    assert ft.RealFourier(fl([1.0, 2.0, 3.0]), G.make_slice(0j, 3, None, 'c'), fft.SCALING_DEFAULT) is not None

    osm = loader.load("oversampling")
    g = gold["oversampling/oversampling_test.go"]
    for factor, fn in ((2, "TestTwoTimesOversampling"), (4, "TestFourTimesOversampling")):
        t = g[fn]
        osd = osm.CreateOversamplerDecimator(factor)
        for x, up, down in zip(t["in"], t["oversampledExpected"], t["decimatedExpected"]):
            u = G.make_slice(0.0, factor * len(x), None, 'f')
            assert osd.Oversample(fl(x), u) is None
            close(u.tolist(), up, 1e-7)
            d = G.make_slice(0.0, len(x), None, 'f')
            assert osd.Decimate(u, d) is None
            close(d.tolist(), down, 1e-7); n += 1

    rs = loader.load("resample")
    g = gold["resample/resample_test.go"]
    t = g["TestTimeSeries"]
    for x, up, down in zip(t["in"], t["outExpectedUp"], t["outExpectedDown"]):
        close(rs.Time(fl(x), 96000, 192000).tolist(), up, 1e-7)
        close(rs.Time(fl(x), 96000, 44100).tolist(), down, 1e-7); n += 1
    t = g["TestFrequencySeries"]
    for x, re, im in zip(t["in"], t["outExpectedReal"], t["outExpectedImag"]):
        r = np.array(rs.Frequency(fl([complex(a, b) for a, b in x], 'c'), 4).tolist())
        close(r.real, re, 1e-7); close(r.imag, im, 1e-7); n += 1
    t = g["TestOversample"]
    for x, e in zip(t["in"], t["outExpected"]):
        tgt = G.make_slice(0.0, len(e), None, 'f')
        rs.Oversample(fl(x), tgt, 2)
        close(tgt.tolist(), e, 1e-7); n += 1

    rnd = loader.load("random")
    t = gold["random/random_test.go"]["TestRNG"]
    for seed, exp in zip(t["seeds"], t["expectedOutputs"]):
        r = rnd.CreatePRNG(seed)
        close([r.NextFloat() for _ in range(8)], exp, 1e-8); n += 1

    circ = loader.load("circular")
    t = gold["circular/circular_test.go"]["TestBuffer"]
    i, e = t["in"], t["expected"]
    b = circ.CreateBuffer(5)
    outs = []

    def retr():
        buf = G.make_slice(0.0, 5, None, 'f')
        assert b.Retrieve(buf) is None
        outs.append(buf.tolist())
    b.Enqueue(*i[0]); retr()
    b.Enqueue(*i[1]); retr()
    b.Enqueue(*i[2]); retr()
    b.Enqueue(*i[3]); b.Enqueue(*i[4]); retr()
    b.Enqueue(*i[5]); retr()
    b.Enqueue(*i[6]); retr()
    for k in (7, 8, 9):
        b.Enqueue(*i[k])
    b.Enqueue(i[10][0]); b.Enqueue(i[11][0]); retr()
    for o, x in zip(outs, e):
        assert o == [float(v) for v in x], (o, x)
        n += 1
    assert b.Retrieve(G.make_slice(0.0, 4, None, 'f')) is not None and b.Length() == 5
    # level/level_test.go TestMeters
    lv = loader.load("level")
    t = json.load(open(os.path.join(HERE, "..", "level_goldens.json")))
    fs, ns = t["sample_rate"], t["samples"]
    sig = np.sin((2.0 * np.pi) * (np.arange(ns, dtype=np.float64) / float(fs)))
    m, err = lv.CreateMeter(2, G.lit(["left", "right"]))
    assert err is None
    rows = lambda: G.lit([fl(a * sig) for a in t["amplitudes"]])
    assert m.Process(rows(), fs) is None                       # created disabled: nothing happens
    read = lambda: [[int(m.Analyze(c)[0].Level()) for c in range(2)], [int(m.Analyze(c)[0].Peak()) for c in range(2)]]
    assert read() == [[-200, -200], [-200, -200]]
    m.SetEnabled(True)
    assert m.Process(rows(), fs) is None
    assert read() == [t["enabled"]["level"], t["enabled"]["peak"]], read()
    m.SetEnabled(False)
    assert read() == [t["after_disable"]["level"], t["after_disable"]["peak"]]
    n += 3

    # wave/wave_test.go: the sample conversions of all six formats (the RIFF container stays out: encoding/binary on structs)
    wv = loader.load("wave")
    t = json.load(open(os.path.join(HERE, "..", "wave_goldens.json")))
    for fmt, fn in (("PCM8", "LPCM8"), ("PCM16", "LPCM16"), ("PCM24", "LPCM24"), ("PCM32", "LPCM32"), ("IEEE32", "IEEE32"), ("IEEE64", "IEEE64")):
        g = t[fmt]
        data, err = getattr(wv, "samplesToBytes" + fn)(fl(g["export_samples"]))
        assert err is None and data.tolist() == g["export_data"], fmt          # TestExport<fmt>Mono: byte-identical payload
        smp, err = getattr(wv, "bytesToSamples" + fn)(G.lit(list(g["import_data"]), "byte"))
        assert err is None
        close(smp.tolist(), g["import_samples"], g["import_tolerance"])          # TestImport<fmt>Mono
        n += 2

    # tuner/tuner_test.go: one of the six recordings (the last 96000 frames: tests/golden/tuner_tail_*.npz)
    if os.environ.get("GO2PY_SELFCHECK_TUNER", "1") != "0":
        tn = loader.load("tuner")
        z = np.load(os.path.join(HERE, "..", "tuner_tail_A2.npz"))
        keys = list(z.keys())
        samples = z["samples"] if "samples" in keys else z[keys[0]]
        tu = tn.Create()
        tu.Process(fl(samples), int(z["sample_rate"]) if "sample_rate" in keys else 96000)
        res, err = tu.Analyze()
        assert err is None, err
        assert res.Note() == "A2" and abs(int(res.Cents())) <= 5 and np.isfinite(res.Frequency()), (res.Note(), res.Cents(), res.Frequency())
        n += 1

    if verbose:
        print("go2py self-check: %d reference test vectors reproduced by the translated packages" % n)
    return n


if __name__ == "__main__":
    import sys
    sys.path.insert(0, HERE)
    import load
    run(load.Loader(sys.argv[1] if len(sys.argv) > 1 else "/root/reference"), verbose=True)
    from goruntime import G
    print("integer-division sites:", dict(G.intdiv_sites))
