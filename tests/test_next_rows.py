"""SURVEY.md 8(f) rows built so far: the wave sample codecs fused into the chain kernel's load/store and
resample.Time on the device - same parity bar as the hot path (bit-exact for the integer/byte work)."""
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


def test_encoded_rejects_unknown_format_and_long_chains(lib):
    c = gdg.Chain(lib=lib)
    import ctypes as C
    buf = (C.c_char * 64)()
    ip = (C.c_void_p * 1)(C.addressof(buf)); op = (C.c_void_p * 1)(C.addressof(buf)); hs = (C.c_void_p * 1)(c._h)
    assert lib.gdg_process_batch_encoded(hs, 1, ip, 9, op, 1, 8, 48000) < 0
    assert b"Unknown sample format" in lib.gdg_last_error()
    assert lib.gdg_format_bytes_per_sample(2) == 3 and lib.gdg_format_bytes_per_sample(7) == 0
    for _ in range(25):
        i = c.append_unit("distortion"); c.set_bypass(i, False)
    with pytest.raises(gdg.GdgError, match="one fused segment"):
        gdg.process_batch_encoded([c], [bytes(16)], "PCM16", "PCM16", 8, 48000)
    # ... while the float64 entry point handles 25 active units in two segments
    a = common.build_chain(orc.Chain(), [("distortion", {}, {})] * 25)
    x = common.synth_fast(0, 48000, 1024) * 3
    assert np.array_equal(c.process(x, 48000), a.process(x, 48000))


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
