"""SURVEY.md 8(f) rows built so far: the wave sample codecs fused into the chain kernel's load/store and
resample.Time on the device - same parity bar as the hot path (bit-exact for the integer/byte work)."""
import ctypes as C_
import json
import os

import numpy as np
import pytest

import common
from common import gdg, orc

BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]
FORMATS = ["PCM8", "PCM16", "PCM24", "PCM32", "IEEE32", "IEEE64"]


@pytest.fixture(params=BACKENDS)
def lib(request):
    return common.backend_lib(request.param)


def _passthrough(lib):
    return [gdg.Chain(lib=lib)]


@pytest.mark.parametrize("fmt", FORMATS)
def test_codecs_match_reference_vectors(lib, fmt):
    """wave/wave_test.go vectors through the device codecs (empty chain = identity):
    import bytes -> float64 (IEEE64 out) and samples (IEEE64 in) -> export bytes, byte-exact."""
    g = json.load(open(os.path.join(common.GOLD, "wave_goldens.json")))[fmt]
    n = 32                                                    # 20 reference samples padded to a power of two
    pad = lambda b, k: bytes(b) + bytes(k)
    bps = orc.wave_bytes_per_sample(fmt)
    # decode
    out = gdg.process_batch_encoded(_passthrough(lib), [pad(g["import_data"], (n - 20) * bps)], fmt, "IEEE64", n, 96000)[0]
    dec = np.frombuffer(out, dtype="<f8")[:20]
    assert np.array_equal(dec, orc.wave_decode(fmt, bytes(g["import_data"])))
    assert np.max(np.abs(dec - np.array(g["import_samples"]))) <= g["import_tolerance"]
    # encode
    x = np.zeros(n); x[:20] = g["export_samples"]
    enc = gdg.process_batch_encoded(_passthrough(lib), [x.tobytes()], "IEEE64", fmt, n, 96000)[0]
    assert list(enc[:20 * bps]) == g["export_data"]


@pytest.mark.parametrize("fmt", FORMATS)
def test_codecs_bit_exact_on_random_and_edge_values(lib, fmt):
    n = 4096
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-1.3, 1.3, n - 16), [0.0, -0.0, 1.0, -1.0, 1.0000001, -1.0000001, 0.5, -0.5,
                                                          1e-9, -1e-9, 0.999999999, -0.999999999, 2.0, -2.0, 1 / 3, -1 / 3]])
    enc = gdg.process_batch_encoded(_passthrough(lib), [x.tobytes()], "IEEE64", fmt, n, 48000)[0]
    assert enc == orc.wave_encode(fmt, x).tobytes()
    dec = gdg.process_batch_encoded(_passthrough(lib), [enc], fmt, "IEEE64", n, 48000)[0]
    assert np.array_equal(np.frombuffer(dec, dtype="<f8"), orc.wave_decode(fmt, enc))
    # all 65536 PCM16 codes round-trip exactly like the reference codec does
    if fmt == "PCM16":
        codes = np.arange(-32768, 32768, dtype="<i2")
        for k in range(0, 65536, 8192):
            blk = codes[k:k + 8192].tobytes()
            rt = gdg.process_batch_encoded(_passthrough(lib), [blk], "PCM16", "PCM16", 8192, 48000)[0]
            assert rt == orc.wave_encode("PCM16", orc.wave_decode("PCM16", blk)).tobytes()


def test_batch_file_path_pcm16_in_pcm24_out(lib):
    """Config 1 shape through the fused codecs: PCM16 payload in, chain, PCM24 payload out, 3 channels x 4 blocks."""
    fs, n, Cn = 48000, 8192, 3
    spec = [("overdrive", {"gain": 12}, {}), ("tone_stack", {}, {})]
    pairs = [common.make_chain_pair(lib, spec) for _ in range(Cn)]
    rng = np.random.default_rng(11)
    pcm = (rng.standard_normal((Cn, 4 * n)) * 5000).clip(-32768, 32767).astype("<i2")
    mism = total = 0
    for k in range(4):
        rows = [pcm[c, k * n:(k + 1) * n].tobytes() for c in range(Cn)]
        got = gdg.process_batch_encoded([p[1] for p in pairs], rows, "PCM16", "PCM24", n, fs)
        for c in range(Cn):
            ref = orc.wave_encode("PCM24", pairs[c][0].process(orc.wave_decode("PCM16", rows[c]), fs)).tobytes()
            a = np.frombuffer(got[c], dtype=np.uint8).reshape(-1, 3).astype(np.int64)
            b = np.frombuffer(ref, dtype=np.uint8).reshape(-1, 3).astype(np.int64)
            va, vb = a[:, 0] | (a[:, 1] << 8) | (a[:, 2] << 16), b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            d = np.abs(((va - vb + (1 << 23)) % (1 << 24)) - (1 << 23))
            assert d.max() <= 1                              # a last-ulp float64 difference can flip one LSB at a truncation boundary
            mism += np.count_nonzero(d); total += len(d)
    assert mism <= total // 2000


def test_encoded_rejects_unknown_format_and_runs_long_chains(lib):
    c = gdg.Chain(lib=lib)
    import ctypes as C
    buf = (C.c_char * 64)()
    ip = (C.c_void_p * 1)(C.addressof(buf)); op = (C.c_void_p * 1)(C.addressof(buf)); hs = (C.c_void_p * 1)(c._h)
    assert lib.gdg_process_batch_encoded(hs, 1, ip, 9, op, 1, 8, 48000) < 0
    assert b"Unknown sample format" in lib.gdg_last_error()
    assert lib.gdg_format_bytes_per_sample(2) == 3 and lib.gdg_format_bytes_per_sample(7) == 0
    for _ in range(25):
        i = c.append_unit("distortion"); c.set_bypass(i, False)
    # 25 active units = two segments: float64 and encoded entry points both handle it (float64 rows between the segments)
    a = common.build_chain(orc.Chain(), [("distortion", {}, {})] * 25)
    x = common.synth_fast(0, 48000, 1024) * 3
    assert np.array_equal(c.process(x, 48000), a.process(x, 48000))
    pcm = orc.wave_encode("PCM16", x).tobytes()
    a2 = common.build_chain(orc.Chain(), [("distortion", {}, {})] * 25)
    got = gdg.process_batch_encoded([c], [pcm], "PCM16", "PCM16", 1024, 48000)[0]
    assert got == orc.wave_encode("PCM16", a2.process(orc.wave_decode("PCM16", pcm), 48000)).tobytes()


# ---- resample.Time -------------------------------------------------------------------------------------
def test_resample_time_reference_vectors(lib, goldens):
    g = goldens["resample/resample_test.go"]["TestTimeSeries"]
    for x, up, down in zip(g["in"], g["outExpectedUp"], g["outExpectedDown"]):
        assert np.max(np.abs(gdg.resample_time(x, 96000, 192000, lib) - up)) <= 1e-7
        assert np.max(np.abs(gdg.resample_time(x, 96000, 44100, lib) - down)) <= 1e-7


@pytest.mark.parametrize("src,dst", [(44100, 192000), (96000, 192000), (192000, 48000), (48000, 48000), (22050, 96000), (88200, 44100)])
def test_resample_time_matches_oracle(lib, src, dst):
    x = common.synth_fast(3, src, 20000)
    a, b = orc.resample_time(x, src, dst), gdg.resample_time(x, src, dst, lib)
    assert len(a) == len(b) == lib.gdg_resample_time_length(len(x), src, dst)
    assert common.rms(a - b) <= common.TOL_RMS
    assert len(gdg.resample_time(np.zeros(0), src, dst, lib)) == 0


# ---- level.Meter -----------------------------------------------------------
METER_RTOL = 1e-10          # state (currentValue / peakValue) relative tolerance: powers by squaring vs. successive products


def _meter_pair(lib, channels):
    m, o = gdg.Meter(channels, lib=lib), orc.Meter(channels)
    assert not m.enabled() and not o.enabled                 # level.go:383-386
    return m, o


def _assert_meters_agree(m, o, channels, exact=False):
    for c in range(channels):
        (mc, mp, mq), (oc, op, oq) = m.state(c), o.state(c)
        assert mq == oq, (c, mq, oq)
        if exact:
            assert (mc, mp) == (oc, op)
        else:
            assert abs(mc - oc) <= METER_RTOL * abs(oc) and abs(mp - op) <= METER_RTOL * abs(op), (c, mc, oc, mp, op)
        assert m.analyze(c) == o.analyze(c)


def test_meter_reference_vector(lib):
    """level/level_test.go TestMeters: one second of a 1 Hz sine in one Process call."""
    g = json.load(open(os.path.join(common.GOLD, "level_goldens.json")))
    fs, n = g["sample_rate"], g["samples"]
    t = np.arange(n, dtype=np.float64) / float(fs)
    sig = np.sin((2.0 * np.pi) * t)
    m = gdg.Meter(2, ["channel_a", "channel_b"], lib=lib)
    assert m.channel_count() == 2 and m.channel_name(0) == "channel_a" and m.channel_name(1) == "channel_b"
    with pytest.raises(gdg.GdgError, match="level meter only has 2 channels"):
        m.channel_name(2)
    m.set_enabled(True)
    m.process([a * sig for a in g["amplitudes"]], fs)
    for c in range(2):
        assert m.analyze(c) == (g["enabled"]["level"][c], g["enabled"]["peak"][c])
    m.set_enabled(False)                                      # clears the state (level.go:226-230)
    for c in range(2):
        assert m.analyze(c) == (g["after_disable"]["level"][c], g["after_disable"]["peak"][c])


@pytest.mark.parametrize("fs", [22050, 96000, 192000])
def test_meter_streams_like_the_reference(lib, fs):
    """blocks of many lengths (one sample .. several segments), bursts, ties, silence, and decays that run past the
    2 s hold time: state and read-outs follow the oracle's literal per-sample loop."""
    rng = np.random.default_rng(fs)
    ch = 4
    m, o = _meter_pair(lib, ch)
    m.set_enabled(True); o.set_enabled(True)
    total = 0
    for k, n in enumerate([8192, 1, 64, 1000, 8192, 8193, 20000, 4096, 512] + [8192] * (3 * fs // 8192 + 2)):
        t = (total + np.arange(n)) / fs
        env = np.exp(-3.0 * t)                                 # decays through the hold time
        rows = [env * rng.uniform(-1, 1, n),
                0.25 * np.sign(np.sin(2 * np.pi * 50 * t)),    # constant magnitude: every sample ties the peak (>=)
                np.zeros(n) if k % 3 else 1e-3 * rng.standard_normal(n),
                (env * np.sin(2 * np.pi * 440 * t)) * (1.0 if k < 6 else 0.0)]
        if k == 4:
            rows[3] = rows[3].copy(); rows[3][n // 2] = 0.9    # a new maximum in the middle of a block
        m.process(rows, fs); o.process(rows, fs)
        total += n
        if k < 12 or k % 5 == 0:
            _assert_meters_agree(m, o, ch)
    _assert_meters_agree(m, o, ch)
    assert o.state(3)[2] > 2 * fs                            # the hold time did expire in the test


def test_meter_enable_disable_and_argument_errors(lib):
    rng = np.random.default_rng(3)
    m, o = _meter_pair(lib, 3)
    rows = [rng.uniform(-1, 1, 2048) for _ in range(3)]
    m.process(rows, 48000); o.process(rows, 48000)           # disabled: nothing happens
    _assert_meters_agree(m, o, 3, exact=True)
    assert m.state(0) == (0.0, 0.0, 0) and m.analyze(0) == (-200, -200)
    m.set_enabled(True); o.set_enabled(True)
    assert m.enabled()
    rows[1] = None                                            # a nil buffer: that channel sees no samples
    m.process(rows, 48000); o.process(rows, 48000)
    _assert_meters_agree(m, o, 3)
    assert m.state(1) == (0.0, 0.0, 0)
    with pytest.raises(gdg.GdgError, match=r"Number of input buffers \(2\) does not match number of channels \(3\)"):
        m.process(rows[:2], 48000)
    with pytest.raises(gdg.GdgError, match="only has 3 channels"):
        m.analyze(3)
    m.set_enabled(False); o.set_enabled(False)
    _assert_meters_agree(m, o, 3, exact=True)
    with pytest.raises(gdg.GdgError, match="provided 1 channel names"):
        gdg.Meter(2, ["only_one"], lib=lib)


def test_meter_low_rate_literal_path(lib):
    """a segment longer than the hold time (sample rates below 4096 Hz) is walked sample by sample: bit-exact"""
    rng = np.random.default_rng(11)
    m, o = _meter_pair(lib, 2)
    m.set_enabled(True); o.set_enabled(True)
    for n in [8192, 3000, 8192]:
        rows = [np.exp(-np.arange(n) / 500.0) * rng.uniform(-1, 1, n), rng.uniform(-1, 1, n) * (np.arange(n) < 10)]
        m.process(rows, 1000); o.process(rows, 1000)
    _assert_meters_agree(m, o, 2, exact=True)
    assert o.state(1)[2] > 2000


def test_controller_process_feeds_the_meter(lib):
    """controller.go:2710-2781: inputs, chain outputs and the master pair reach the meter (port layout 2C+3)."""
    fs, n, C = 48000, 1024, 5
    gc, oc = [], []
    for c in range(C):
        a, b = common.make_chain_pair(lib, [("overdrive", {"gain": 5 + c}, {}), ("delay", {"delay_time": 10}, {})])
        oc.append(a); gc.append(b)
    gs, os_ = gdg.Spatializer(C, lib=lib), orc.Spatializer(C)
    for c in range(C):
        gs.set_azimuth(c, -60.0 + 30.0 * c); os_.set_azimuth(c, -60.0 + 30.0 * c)
    m, o = _meter_pair(lib, 2 * C + 3)
    m.set_enabled(True); o.set_enabled(True)
    for blk in range(3):
        x = np.stack([common.synth_fast(c, fs, 4 * n, 1)[blk * n:(blk + 1) * n] for c in range(C)])
        got = gdg.controller_process(gc, None, -1, gs, x, fs, meter=m)
        ref = orc.controller_process(oc, None, -1, os_, x, fs)
        assert np.max(np.abs(got - ref)) < 1e-9
        o.process(list(x) + list(ref[:C]) + [None] + list(ref[C:]), fs)
    _assert_meters_agree(m, o, 2 * C + 3)
    assert m.analyze(0)[0] > -60 and m.analyze(2 * C + 1)[0] > -60 and m.analyze(2 * C) == (-200, -200)
    with pytest.raises(gdg.GdgError, match="level meter has 4 channels"):
        gdg.controller_process(gc, None, -1, gs, x, fs, meter=gdg.Meter(4, lib=lib))


# ---- submit / wait: two blocks in flight --------------------------------------------------------------
def test_controller_submit_wait_pipeline_matches_sequential(lib):
    """gdg_controller_submit / gdg_controller_wait: block k+1 is submitted before block k is waited for; outputs, tuner
    and meter state are those of the sequential reference loop (controller.go:2991-3106 walks blocks in order)."""
    fs, n, C, blocks = 96000, 2048, 6, 6
    spec = [("compressor", {}, {}), ("overdrive", {"gain": 12}, {"oversampling": "2"}), ("chorus", {}, {}), ("delay", {"delay_time": 30}, {})]
    oc, gc = zip(*[common.make_chain_pair(lib, spec) for _ in range(C)])
    gs, os_ = gdg.Spatializer(C, lib=lib), orc.Spatializer(C)
    gt, ot = gdg.Tuner(lib=lib), orc.Tuner()
    m, o = _meter_pair(lib, 2 * C + 3)
    m.set_enabled(True); o.set_enabled(True)
    x = np.stack([common.synth_fast(c, fs, blocks * n, 2) for c in range(C)])
    xin = [np.ascontiguousarray(x[:, k * n:(k + 1) * n]) for k in range(blocks)]
    outs = [np.zeros((C + 2, n)) for _ in range(blocks)]
    tickets = []
    for k in range(blocks):
        tickets.append(gdg.controller_submit(list(gc), gt, 1, gs, xin[k], outs[k], fs, meter=m))
        if k >= 1:
            gdg.controller_wait(tickets[k - 1], lib)
    with pytest.raises(gdg.GdgError, match="not in flight"):
        gdg.controller_wait(tickets[0], lib)                  # already waited for
    gdg.controller_wait(tickets[-1], lib)
    for k in range(blocks):
        ref = orc.controller_process(list(oc), ot, 1, os_, xin[k], fs)
        o.process(list(xin[k]) + list(ref[:C]) + [None] + list(ref[C:]), fs)
        for c in range(C + 2):
            assert common.rms(ref[c] - outs[k][c]) <= common.TOL_RMS, (k, c)
    _assert_meters_agree(m, o, 2 * C + 3)
    assert gt.analyze()["note"] == ot.analyze()["note"]
    # a third submit without a wait is refused, and the synchronous call still works afterwards
    t0 = gdg.controller_submit(list(gc), None, -1, gs, xin[0], outs[0], fs)
    t1 = gdg.controller_submit(list(gc), None, -1, gs, xin[1], outs[1], fs)
    assert {t0, t1} == {0, 1}
    with pytest.raises(gdg.GdgError, match="two submitted blocks are in flight"):
        gdg.controller_submit(list(gc), None, -1, gs, xin[2], outs[2], fs)
    with pytest.raises(gdg.GdgError, match="two submitted blocks are in flight"):
        gdg.controller_process(list(gc), None, -1, gs, xin[2], fs)
    gdg.controller_wait(t0, lib); gdg.controller_wait(t1, lib)
    assert gdg.controller_process(list(gc), None, -1, gs, xin[2], fs).shape == (C + 2, n)


def test_encoded_submit_wait(lib):
    """gdg_process_batch_encoded_submit: PCM16 blocks with two in flight give the bytes of the blocking call."""
    import ctypes as Ct
    fs, n, C, blocks = 48000, 1024, 3, 4
    spec = [("overdrive", {"gain": 9}, {"oversampling": "2"}), ("tone_stack", {}, {}), ("delay", {"delay_time": 5}, {})]
    ref_chains = [common.make_chain_pair(lib, spec)[1] for _ in range(C)]
    chains = [common.make_chain_pair(lib, spec)[1] for _ in range(C)]
    x = [(np.stack([common.synth_fast(c, fs, blocks * n, 7)[k * n:(k + 1) * n] for c in range(C)]) * 20000).astype("<i2") for k in range(blocks)]
    want = [gdg.process_batch_encoded(ref_chains, [r.tobytes() for r in x[k]], "PCM16", "PCM16", n, fs) for k in range(blocks)]
    outs = [np.zeros((C, n), dtype="<i2") for _ in range(blocks)]
    handles = (Ct.c_void_p * C)(*[c._h for c in chains])
    tickets = []
    for k in range(blocks):
        ip = (Ct.c_void_p * C)(*[x[k][c].ctypes.data for c in range(C)])
        op = (Ct.c_void_p * C)(*[outs[k][c].ctypes.data for c in range(C)])
        t = lib.gdg_process_batch_encoded_submit(handles, C, ip, 1, op, 1, n, fs)
        assert t >= 0, lib.gdg_last_error()
        tickets.append(t)
        if k >= 1:
            assert lib.gdg_controller_wait(tickets[k - 1]) == 0
    assert lib.gdg_controller_wait(tickets[-1]) == 0
    for k in range(blocks):
        for c in range(C):
            assert outs[k][c].tobytes() == bytes(want[k][c]), (k, c)


@pytest.mark.parametrize("fmt,aux", [("PCM16", "tsm"), ("PCM24", "tsm"), ("IEEE32", "tsm"), ("PCM16", "s"), ("PCM24", "ts"), ("PCM8", "t"),
                                     ("PCM32", "sm")])
def test_controller_submit_encoded(lib, fmt, aux):
    """gdg_controller_submit_encoded: encoded rows in, encoded rows (channels + master pair) out, tuner / spatializer /
    meter on the float64 in between - byte for byte what decoding, the float64 controller call and encoding give.
    aux: which of tuner / spatializer / meter are present (the meter decides whether the input payload is converted ahead of
    the chain kernels, the spatializer / meter whether the last segment keeps a float64 copy of its result)."""
    import ctypes as Ct
    fs, n, C, blocks = 96000, 2048, 4, 3
    spec = [("overdrive", {"gain": 9}, {"oversampling": "2"}), ("chorus", {}, {}), ("reverb", {"mix": 30}, {})]
    ref_chains = [common.make_chain_pair(lib, spec)[1] for _ in range(C)]
    chains = [common.make_chain_pair(lib, spec)[1] for _ in range(C)]
    rs, s = gdg.Spatializer(C, lib=lib), gdg.Spatializer(C, lib=lib)
    rt, t = gdg.Tuner(lib=lib), gdg.Tuner(lib=lib)
    rm, m = gdg.Meter(2 * C + 3, lib=lib), gdg.Meter(2 * C + 3, lib=lib)
    rm.set_enabled(True); m.set_enabled(True)
    for c in range(C):
        rs.set_azimuth(c, -45.0 + 30.0 * c); s.set_azimuth(c, -45.0 + 30.0 * c)
    if "t" not in aux: rt = t = None
    if "s" not in aux: rs = s = None
    if "m" not in aux: rm = m = None
    R = C + (2 if s else 0)
    fi = gdg.FORMATS.index(fmt)
    bps = orc.wave_bytes_per_sample(fmt)
    handles = (Ct.c_void_p * C)(*[c._h for c in chains])
    x = np.stack([common.synth_fast(c, fs, blocks * n, 9) for c in range(C)]) * 0.8
    pay = [[np.frombuffer(orc.wave_encode(fmt, x[c, k * n:(k + 1) * n]).tobytes(), dtype=np.uint8).copy() for c in range(C)] for k in range(blocks)]
    outs = [[np.zeros(n * bps, dtype=np.uint8) for _ in range(R)] for _ in range(blocks)]
    prev = None
    for k in range(blocks):
        ip = (Ct.c_void_p * C)(*[r.ctypes.data for r in pay[k]])
        op = (Ct.c_void_p * R)(*[r.ctypes.data for r in outs[k]])
        tk = lib.gdg_controller_submit_encoded(handles, C, t._h if t else None, 2, s._h if s else None, m._h if m else None, ip, fi, op, fi, R, n, fs)
        assert tk >= 0, lib.gdg_last_error()
        if prev is not None:
            assert lib.gdg_controller_wait(prev) == 0
        prev = tk
    assert lib.gdg_controller_wait(prev) == 0
    for k in range(blocks):
        dec = np.stack([orc.wave_decode(fmt, pay[k][c].tobytes()) for c in range(C)])
        y = gdg.controller_process(ref_chains, rt, 2, rs, dec, fs, meter=rm)
        for c in range(R):
            assert outs[k][c].tobytes() == orc.wave_encode(fmt, y[c]).tobytes(), (k, c)
    if t: assert t.analyze() == rt.analyze()
    for c in (0, C, 2 * C + 1, 2 * C + 2) if m else ():
        assert m.state(c) == rm.state(c)


# ---- the metronome row (auxInputBuffer) through the controller mirror, setters under load ----------------------
def test_controller_process_metronome_into_master_and_meter(lib):
    """controller.go:2722-2761: the metronome's buffer feeds meter port 2C and - with metrMasterOutput - both master channels
    through spatializer.Process' auxInputBuffer (spatializer.go:300-308)."""
    fs, n, C = 48000, 512, 3
    gc, oc = [], []
    for c in range(C):
        a, b = common.make_chain_pair(lib, [("distortion", {"gain": 6 + c}, {}), ("tremolo", {}, {})])
        oc.append(a); gc.append(b)
    gs, os_ = gdg.Spatializer(C, lib=lib), orc.Spatializer(C)
    for c in range(C):
        gs.set_azimuth(c, -45.0 + 45.0 * c); os_.set_azimuth(c, -45.0 + 45.0 * c)
    m, o = _meter_pair(lib, 2 * C + 3)
    m.set_enabled(True); o.set_enabled(True)
    rng = np.random.default_rng(4)
    for blk in range(4):
        x = np.stack([common.synth_fast(c, fs, 4 * n, 3)[blk * n:(blk + 1) * n] for c in range(C)])
        click = np.where(np.arange(n) < 40, 0.25 * rng.uniform(-1, 1, n), 0.0)
        to_master = blk % 2 == 0
        got = gdg.controller_process(gc, None, -1, gs, x, fs, meter=m, metronome=click, metronome_to_master=to_master)
        outs = np.stack([oc[c].process(x[c], fs) for c in range(C)])
        master = os_.process(outs, click if to_master else None)
        assert np.max(np.abs(got[:C] - outs)) < 1e-9 and np.max(np.abs(got[C:] - master)) < 1e-9
        o.process(list(x) + list(outs) + [click] + list(master), fs)
    _assert_meters_agree(m, o, 2 * C + 3)
    assert m.analyze(2 * C)[0] > -60                         # the metronome port saw the click


def test_tuner_is_fed_even_without_output_rows(lib):
    """controller.go:2666-2672 feeds the tuner before it looks at the number of outputs; with nOut < nIn nothing else runs."""
    fs, n, C = 48000, 4096, 2
    chains = [common.build_chain(gdg.Chain(lib=lib), [("tone_stack", {}, {})]) for _ in range(C)]
    gt, ot = gdg.Tuner(lib=lib), orc.Tuner()
    t = np.arange(24 * n) / fs
    x = 0.5 * np.sin(2 * np.pi * 110.0 * t)
    for k in range(24):
        rows = np.stack([x[k * n:(k + 1) * n], np.zeros(n)])
        ip = (C_.c_void_p * C)(*[rows[i].ctypes.data for i in range(C)])
        out = np.zeros((1, n))
        op = (C_.c_void_p * 1)(out[0].ctypes.data)
        assert lib.gdg_controller_process(gdg._handles(chains), C, gt._h, 0, None, ip, op, 1, n, fs) == 0
        assert not out.any()
        ot.process(rows[0], fs)
    assert gt.analyze()["note"] == ot.analyze()["note"] == "A2"


def test_setters_race_free_against_processing(lib):
    """effects.go:98 / signal.go:45: parameter and topology edits arrive from another thread while blocks are processed.
    Every block must equal one of the oracle's outputs for SOME interleaving - here the edits toggle between two whole
    configurations under a lock shared with the checker, so each block is checked against the configuration it ran with;
    the point of the test is that nothing crashes, deadlocks or reads freed state (RemoveUnit frees device memory)."""
    import threading
    fs, n = 48000, 256
    chain = common.build_chain(gdg.Chain(lib=lib), [("overdrive", {"gain": 10}, {}), ("delay", {"delay_time": 5}, {}), ("cabinet", {}, {})])
    stop = threading.Event()
    errors = []

    def hammer():
        k = 0
        try:
            while not stop.is_set():
                chain.set_numeric_value(0, "gain", 5 + k % 20)
                chain.set_bypass(1, k % 3 == 0)
                i = chain.append_unit("tremolo")
                chain.set_bypass(i, False)
                chain.get_numeric_value(0, "gain"), chain.parameters(0), chain.length()
                chain.remove_unit(i)
                if k % 5 == 0:
                    chain.move_down(0), chain.move_up(1)
                k += 1
        except Exception as e:                               # noqa: BLE001
            errors.append(e)
    th = threading.Thread(target=hammer)
    th.start()
    x = common.synth_fast(1, fs, 150 * n)
    try:
        for k in range(150):
            y = chain.process(x[k * n:(k + 1) * n], fs)
            assert np.all(np.isfinite(y)) and np.max(np.abs(y)) <= 1.0 + 1e-12
    finally:
        stop.set(); th.join()
    assert not errors, errors


def test_spatializer_setters_from_a_second_thread_are_never_lost(lib):
    """spatializer.go:59: SetAzimuth / SetDistance / SetLevel arrive from the UI thread while Process runs.  Every block must
    be finite and bounded, and once the edits stop the NEXT block must be computed from the final positions - a setter
    that lands between "positions read" and "clean flag set" on the processing side must not be forgotten."""
    import threading
    fs, n, C = 48000, 256, 6
    gs, osp = gdg.Spatializer(C, lib=lib), orc.Spatializer(C)
    gs.set_sample_rate(fs); osp.set_sample_rate(fs)
    x = np.stack([common.synth_fast(c, fs, 64 * n) for c in range(C)])
    final = [(-75.0 + 30.0 * c, 0.5 + 0.7 * c, 1.0 - 0.1 * c) for c in range(C)]
    for rounds in range(12):
        stop = threading.Event()
        errors = []

        def hammer():
            k = 0
            try:
                while not stop.is_set():
                    c = k % C
                    gs.set_azimuth(c, -90.0 + (7 * k) % 180); gs.set_distance(c, (k % 100) / 10.0); gs.set_level(c, (k % 11) / 10.0)
                    gs.get_azimuth(c)
                    k += 1
            except Exception as e:                           # noqa: BLE001
                errors.append(e)
            for c, (az, d, lv) in enumerate(final):         # the last thing the thread does: the final positions
                gs.set_azimuth(c, az); gs.set_distance(c, d); gs.set_level(c, lv)
        th = threading.Thread(target=hammer)
        th.start()
        try:
            for k in range(4):
                y = gs.process(x[:, (4 * rounds + k) * n:(4 * rounds + k + 1) * n])
                assert np.all(np.isfinite(y)) and np.max(np.abs(y)) <= C + 1e-9
        finally:
            stop.set(); th.join()
        assert not errors, errors
    # histories are the inputs themselves (spatializer.go:310-333) and do not depend on the positions: feed the oracle the
    # same stream, set the final positions, and the next block must match
    for k in range(48):
        osp.process(x[:, k * n:(k + 1) * n])
    for c, (az, d, lv) in enumerate(final):
        osp.set_azimuth(c, az); osp.set_distance(c, d); osp.set_level(c, lv)
    blk = x[:, 48 * n:49 * n]
    ya, yb = osp.process(blk), gs.process(blk)
    assert common.rms(ya - yb) <= common.TOL_RMS and common.rms(ya) > 0


def test_tuner_analyze_from_a_second_thread_during_processing(lib):
    """tuner.go:50-53 (mutexBuffer / mutexAnalyze): the UI polls Analyze (controller.go:625-627) while the audio thread
    keeps feeding the ring through controller.process.  A steady tone must read as the same note whenever the poll
    lands, and the final read-out equals the oracle's."""
    import threading
    fs, n, C = 48000, 2048, 2
    chains = [common.build_chain(gdg.Chain(lib=lib), [("tone_stack", {}, {})]) for _ in range(C)]
    gt, ot = gdg.Tuner(lib=lib), orc.Tuner()
    blocks = 96
    t = np.arange(blocks * n) / fs
    x = 0.5 * np.sin(2 * np.pi * 146.83 * t)                 # D3
    for k in range(48):                                      # fill the 96000-sample ring once before polling starts
        gt.process(x[k * n:(k + 1) * n], fs); ot.process(x[k * n:(k + 1) * n], fs)
    stop = threading.Event()
    seen, errors = [], []

    def poll():
        try:
            while not stop.is_set():
                seen.append(gt.analyze()["note"])
                gdg.tuner_analyze_batch([gt])
        except Exception as e:                               # noqa: BLE001
            errors.append(e)
    th = threading.Thread(target=poll)
    th.start()
    try:
        for k in range(48, blocks):
            rows = np.stack([x[k * n:(k + 1) * n], np.zeros(n)])
            gdg.controller_process(chains, gt, 0, None, rows, fs)
            ot.process(rows[0], fs)
    finally:
        stop.set(); th.join()
    assert not errors, errors
    assert seen and set(seen) == {"D3"}, sorted(set(seen))
    a, b = ot.analyze(), gt.analyze()
    assert a["note"] == b["note"] == "D3" and a["cents"] == b["cents"] and abs(a["frequency"] - b["frequency"]) <= 1e-9 * a["frequency"]


def test_getters_reject_null_result_pointers(lib):
    c = common.build_chain(gdg.Chain(lib=lib), [("overdrive", {}, {})])
    assert lib.gdg_chain_get_bypass(c._h, 0, None) < 0
    assert lib.gdg_chain_get_numeric_value(c._h, 0, b"gain", None) < 0 and lib.gdg_chain_get_numeric_value(c._h, 0, None, None) < 0
    assert lib.gdg_chain_get_discrete_value(c._h, 0, b"valve", None) < 0
    assert lib.gdg_chain_set_numeric_value(c._h, 0, None, 3) < 0
    assert lib.gdg_chain_set_discrete_value(c._h, 0, b"valve", None) < 0 and lib.gdg_chain_set_discrete_value(c._h, 0, None, b"x") < 0
    assert lib.gdg_chain_parameter(c._h, 0, 0, None) < 0
    assert b"bad argument" in lib.gdg_last_error()
    s = gdg.Spatializer(2, lib=lib)
    assert lib.gdg_spatializer_get_azimuth(s._h, 0, None) < 0 and lib.gdg_spatializer_get_level(s._h, 1, None) < 0
    assert c.get_numeric_value(0, "gain") == 0 and c.get_discrete_value(0, "valve")      # the chain is untouched


def test_irbank_names_stay_valid_between_adds(lib):
    """ImpulseResponses.Names() (filter.go:681-699): the strings a caller got stay where they are until the bank changes
    (a cgo caller reads them with C.GoString after the call; a second lookup must not move the first one's bytes)"""
    bank = gdg.IRBank(lib)
    for k in range(3):
        bank.add_raw("response %d" % k, 0, common.synthetic_ir(64), 48000)
    raw = C_.CFUNCTYPE(C_.c_void_p, C_.c_void_p, C_.c_size_t)(("gdg_irbank_name", lib))
    first = [raw(bank._h, i) for i in range(3)]
    again = [raw(bank._h, i) for i in reversed(range(3))][::-1]
    assert first == again and all(first)
    assert [C_.string_at(q) for q in first] == [b"response 0", b"response 1", b"response 2"]
    assert raw(bank._h, 3) is None
    bank.add_raw("response 3", 0, common.synthetic_ir(64), 48000)
    assert bank.names() == ["response 0", "response 1", "response 2", "response 3"]


# ---- impulse-response compilation on the device, batched tuner analysis ------------------------------
def _oracle_composite(h_by_name, comp_db, rate_src, rate, order, picks):
    """poweramp.compile (poweramp.go:25-127) with the oracle's filter algebra: picks = [(name, level_db)]"""
    acc = np.zeros(0)
    for name, level in picks:
        data = orc.resample_time(h_by_name[name], rate_src, rate)       # filter.Import resamples to EVERY rate, its own included (filter.go:766-780)
        f = orc.Filter(data, rate)
        f.set_compensation(10.0 ** (0.05 * comp_db[name]))
        if order > 0:
            f = f.reduce(order)
        c = f.normalize().multiply(10.0 ** (0.05 * level)).coefficients()
        out = np.zeros(max(len(acc), len(c)))
        out[:len(acc)] = acc
        out[:len(c)] += c
        acc = out
    return acc


@pytest.mark.parametrize("order", ["1048576", "4096", "512", "64"])
def test_device_ir_compile_matches_oracle_algebra(lib, order):
    """filter.Import's resampling, Reduce, Normalize, Multiply and Add run on the device (gdg_host.cpp compile_poweramp);
    the composite's coefficients equal the oracle's filter algebra on the same responses."""
    rng = np.random.default_rng(17)
    src_rate = 96000
    h = {"A": rng.uniform(-1, 1, 3000) * np.exp(-np.arange(3000) / 500.0), "B": rng.uniform(-1, 1, 700) * np.exp(-np.arange(700) / 90.0),
         "C": rng.uniform(-1, 1, 9600) * np.exp(-np.arange(9600) / 2500.0)}
    comp = {"A": -3, "B": 2, "C": 0}
    bank = gdg.IRBank(lib)
    for k in h:
        bank.add(k, comp[k], h[k], src_rate)
    picks = [("A", -2), ("C", -7), ("B", 0)]
    for rate in (192000, 44100, 96000):
        ch = gdg.Chain(bank, lib=lib)
        i = ch.append_unit("power_amp")
        ch.set_discrete_value(i, "filter_order", order)
        for slot, (name, level) in enumerate(picks):
            ch.set_discrete_value(i, "filter_%d" % (slot + 1), name)
            ch.set_numeric_value(i, "level_%d" % (slot + 1), level)
        ch.set_bypass(i, False)
        assert len(ch.filter_coefficients(i)) == 0               # no rate yet: nothing compiled (poweramp.go:132-154 with sampleRate 0)
        ch.process(np.zeros(256), rate)
        got = ch.filter_coefficients(i)
        ref = _oracle_composite(h, comp, src_rate, rate, int(order), picks)
        assert got.shape == ref.shape, (rate, got.shape, ref.shape)
        assert np.max(np.abs(got - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref))), (rate, np.max(np.abs(got - ref)))


def test_identical_recipes_share_one_compiled_filter(lib):
    """1024 chains with the same power-amp settings compile ONE composite filter and ONE spectrum (the recipe cache)."""
    bank = gdg.IRBank(lib)
    bank.add_raw("ir", 0, common.synthetic_ir(300), 48000)
    before = gdg.launch_count(lib)
    chains = []
    for _ in range(64):
        c = common.build_chain(gdg.Chain(bank, lib=lib), [("power_amp", {"level_1": -3}, {"filter_1": "ir"})])
        chains.append(c)
    x = np.zeros((64, 256))
    gdg.process_batch(chains, x, 48000)
    launches = gdg.launch_count(lib) - before
    assert launches <= 8, launches                               # sum of squares, scale-add, spectra, the chain kernel - not 64 of each


def test_tuner_analyze_batch_equals_single_calls(lib):
    import os
    notes = ["A2", "D3", "E4"]
    tuners, singles = [], []
    for note in notes:
        z = np.load(os.path.join(common.GOLD, "tuner_tail_%s.npz" % note))
        t = gdg.Tuner(lib=lib)
        t.process(z["samples"], int(z["sample_rate"]))
        tuners.append(t)
        o = orc.Tuner()
        o.process(z["samples"], int(z["sample_rate"]))
        singles.append(o.analyze())
    batch = gdg.tuner_analyze_batch(tuners)
    for note, b, o in zip(notes, batch, singles):
        assert b["note"] == note == o["note"] and b["cents"] == o["cents"]
        assert abs(b["frequency"] - o["frequency"]) <= 1e-9 * o["frequency"]
    assert [t.analyze() for t in tuners] == batch


# ---- ABI entry points that had no test of their own ---------------------------------------------------------------
def test_controller_submit_aux_two_blocks_in_flight(lib):
    """gdg_controller_submit_aux: the ticketed call with the metronome row (controller.go:2722-2761), two blocks in flight,
    against the blocking call on a second set of chains."""
    fs, n, C, blocks = 96000, 1024, 4, 6
    spec = [("overdrive", {"gain": 9}, {"oversampling": "2"}), ("delay", {"delay_time": 5}, {})]
    a = [common.build_chain(gdg.Chain(lib=lib), spec) for _ in range(C)]
    b = [common.build_chain(gdg.Chain(lib=lib), spec) for _ in range(C)]
    sa, sb = gdg.Spatializer(C, lib=lib), gdg.Spatializer(C, lib=lib)
    ma, mb = gdg.Meter(2 * C + 3, lib=lib), gdg.Meter(2 * C + 3, lib=lib)
    ma.set_enabled(True); mb.set_enabled(True)
    for c in range(C):
        sa.set_azimuth(c, -60.0 + 40.0 * c); sb.set_azimuth(c, -60.0 + 40.0 * c)
    rng = np.random.default_rng(23)
    x = np.stack([common.synth_fast(c, fs, blocks * n, 2) for c in range(C)])
    ins = [np.ascontiguousarray(x[:, k * n:(k + 1) * n]) for k in range(blocks)]
    clicks = [np.where(np.arange(n) < 64, 0.2 * rng.uniform(-1, 1, n), 0.0) for _ in range(blocks)]
    outs = [np.zeros((C + 2, n)) for _ in range(blocks)]
    prev = None
    for k in range(blocks):
        t = gdg.controller_submit(a, None, -1, sa, ins[k], outs[k], fs, meter=ma, metronome=clicks[k], metronome_to_master=k % 3 != 1)
        if prev is not None:
            gdg.controller_wait(prev, lib)
        prev = t
    gdg.controller_wait(prev, lib)
    for k in range(blocks):
        ref = gdg.controller_process(b, None, -1, sb, ins[k], fs, meter=mb, metronome=clicks[k], metronome_to_master=k % 3 != 1)
        assert np.array_equal(outs[k], ref), k
        assert np.max(np.abs(ref[C:])) > 0
    for c in range(2 * C + 3):
        assert ma.state(c) == mb.state(c), c
    assert ma.analyze(2 * C)[0] > -60


def _device_rows(lib, host_u8):
    """a [rows][bytes] uint8 array as device memory of the backend under test: (pointer, keep-alive, fetch())"""
    if gdg.backend_name(lib).startswith("cuda"):
        import torch
        t = torch.from_numpy(host_u8).cuda()
        return t.data_ptr(), t, lambda: t.cpu().numpy()
    buf = host_u8.copy()                                   # the emulation's device memory is host memory
    return buf.ctypes.data, buf, lambda: buf


@pytest.mark.parametrize("fmt_in,fmt_out", [("PCM16", "PCM24"), ("IEEE32", "PCM16"), ("PCM8", "IEEE64")])
def test_process_batch_device_encoded(lib, fmt_in, fmt_out):
    """gdg_process_batch_device_encoded: payload rows that already sit on the device (row strides larger than the rows),
    byte-exact against the host-row call on a second set of chains."""
    fs, n, C, blocks = 48000, 2048, 5, 3
    spec = [("distortion", {"gain": 12}, {"oversampling": "4"}), ("tone_stack", {}, {}), ("chorus", {}, {})]
    a = [common.build_chain(gdg.Chain(lib=lib), spec) for _ in range(C)]
    b = [common.build_chain(gdg.Chain(lib=lib), spec) for _ in range(C)]
    runner = gdg.BatchRunner(a)
    bi = lib.gdg_format_bytes_per_sample(gdg.FORMATS.index(fmt_in)); bo = lib.gdg_format_bytes_per_sample(gdg.FORMATS.index(fmt_out))
    si, so = n * bi + 64, n * bo + 128                     # strides: rows do not touch
    x = np.stack([common.synth_fast(c, fs, blocks * n, 2) for c in range(C)])
    for k in range(blocks):
        rows = [orc.wave_encode(fmt_in, x[c, k * n:(k + 1) * n]).tobytes() for c in range(C)]
        hin = np.zeros((C, si), dtype=np.uint8)
        for c in range(C):
            hin[c, :n * bi] = np.frombuffer(rows[c], dtype=np.uint8)
        pin, keep_in, _ = _device_rows(lib, hin)
        pout, keep_out, fetch = _device_rows(lib, np.full((C, so), 0xA5, dtype=np.uint8))
        runner.process_device_encoded(pin, fmt_in, si, pout, fmt_out, so, n, fs)
        lib.gdg_synchronize()
        got = fetch()
        want = gdg.process_batch_encoded(b, rows, fmt_in, fmt_out, n, fs)
        for c in range(C):
            assert got[c, :n * bo].tobytes() == want[c], (k, c)
            assert np.all(got[c, n * bo:] == 0xA5)          # nothing written between the rows
    with pytest.raises(gdg.GdgError, match="stride too small"):
        runner.process_device_encoded(pin, fmt_in, n * bi - 1, pout, fmt_out, so, n, fs)
    if "IEEE" in fmt_in:                                     # float rows are read as words: a row may not start off the sample grid
        with pytest.raises(gdg.GdgError, match="aligned to the sample size"):
            runner.process_device_encoded(pin, fmt_in, si + 1, pout, fmt_out, so, n, fs)
    if "IEEE" in fmt_out:
        with pytest.raises(gdg.GdgError, match="aligned to the sample size"):
            runner.process_device_encoded(pin, fmt_in, si, pout + 4, fmt_out, so, n, fs)
