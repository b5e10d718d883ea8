"""The batch-file loop (controller.processFiles, controller.go:2985-3107) through go-dsp-guitar_b200/batch.py: inputs of
different lengths and rates -> resample.Time -> zero-padded blocks of 8192 -> controller.process with two blocks in flight ->
encoded output rows, against the same loop written out with the oracle."""
import importlib
import json
import os

import numpy as np
import pytest

import common
from common import gdg, orc

batch = importlib.import_module("go-dsp-guitar_b200.batch")
patch = importlib.import_module("go-dsp-guitar_b200.patch")
BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]
N = 8192
SPECS = [[("overdrive", {"gain": 12}, {}), ("tone_stack", {}, {})],                                   # BASELINE config 1's chain
         [("compressor", {}, {}), ("distortion", {"gain": 8}, {"oversampling": "2"}), ("delay", {"delay_time": 120}, {}), ("cabinet", {}, {})],
         [("chorus", {}, {}), ("reverb", {"mix": 30}, {})]]


@pytest.fixture(params=BACKENDS)
def lib(request):
    return common.backend_lib(request.param)


def _oracle_files(specs, geometry, inputs, rates, target, fmt):
    chains = [common.build_chain(orc.Chain(), sp) for sp in specs]
    spat = orc.Spatializer(len(specs))
    spat.set_sample_rate(target)
    for c, (az, d, lv) in enumerate(geometry):
        spat.set_azimuth(c, az); spat.set_distance(c, d); spat.set_level(c, lv)
    xs = [orc.resample_time(x, r, target) if r != target else np.asarray(x, dtype=np.float64) for x, r in zip(inputs, rates)]
    L = max(len(x) for x in xs)
    if L % N:
        L = N * (L // N + 1)
    x = np.zeros((len(xs), L))
    for c, v in enumerate(xs):
        x[c, :len(v)] = v
    out = np.zeros((len(xs) + 2, L))
    for k in range(L // N):
        sl = slice(k * N, (k + 1) * N)
        out[:len(xs), sl] = np.stack([chains[c].process(x[c, sl], target) for c in range(len(xs))])
        out[len(xs):, sl] = spat.process(out[:len(xs), sl])
    return [orc.wave_encode(fmt, row).tobytes() for row in out] + [orc.wave_encode(fmt, np.zeros(L)).tobytes()], L


@pytest.mark.parametrize("fmt", ["PCM16", "PCM24", "IEEE32"])
def test_batch_loop_matches_the_oracle_loop(lib, fmt):
    target = 48000
    rates = [48000, 44100, 48000]
    lens = [2 * N + 1234, 3 * N - 500, N]                       # different lengths: everything is padded to the longest, whole blocks
    inputs = [common.synth_fast(c, rates[c], lens[c]) for c in range(3)]
    geometry = [(-40.0, 1.0, 1.0), (0.0, 2.0, 0.7), (55.0, 1.5, 0.9)]
    chains = [common.build_chain(gdg.Chain(lib=lib), sp) for sp in SPECS]
    spat = gdg.Spatializer(3, lib=lib)
    spat.set_sample_rate(target)
    for c, (az, d, lv) in enumerate(geometry):
        spat.set_azimuth(c, az); spat.set_distance(c, d); spat.set_level(c, lv)
    got = batch.process_files(chains, spat, inputs, rates, target, out_format=fmt)
    want, L = _oracle_files(SPECS, geometry, inputs, rates, target, fmt)
    assert L == 4 * N and len(got) == len(want) == 3 + batch.MORE_OUTPUTS_THAN_INPUTS
    bps = {"PCM16": 2, "PCM24": 3, "IEEE32": 4}[fmt]
    for r, (g, w) in enumerate(zip(got, want)):
        assert len(g) == len(w) == L * bps, r
        a, b = orc.wave_decode(fmt, g), orc.wave_decode(fmt, w)
        lsb = {"PCM16": 2.0 / 65535.0, "PCM24": 2.0 / 16777215.0, "IEEE32": 1e-7}[fmt]
        assert np.max(np.abs(a - b)) <= 1.000001 * lsb, (r, np.max(np.abs(a - b)))           # a 1e-16 difference may flip one code
        assert np.mean(np.frombuffer(g, dtype=np.uint8) != np.frombuffer(w, dtype=np.uint8)) < 1e-3, r
        assert r == len(got) - 1 or np.abs(b).max() > 1e-3, r                               # every row but the metronome port carries signal
    assert got[-1] == want[-1]                                                              # the metronome port: encoded silence


def test_batch_loop_from_the_patch_file_without_spatializer_and_with_empty_inputs(lib):
    """chains restored from examples/chain12.patch.json (go-dsp-guitar_b200/patch.py); no spatializer: the master rows are silence
    like the reference's untouched output buffers; an empty input (a file that failed to load, controller.go:2911-2913) is a row of zeros"""
    _ob, pb = common.make_banks(lib)
    ob, _ = common.make_banks(lib)
    cfg = json.load(open(os.path.join(common.ROOT, "examples", "chain12.patch.json")))
    two = dict(cfg, Channels=cfg["Channels"][:2])
    chains = [gdg.Chain(pb, lib=lib) for _ in range(2)]
    ochains = [orc.Chain(ob) for _ in range(2)]
    assert patch.restore(two, chains) == [] and patch.restore(two, ochains) == []
    fs = 96000
    x0 = common.synth_fast(0, fs, N + 100)
    got = batch.process_files(chains, None, [x0, np.zeros(0)], [fs, fs], fs, out_format="IEEE64")
    assert len(got) == 5 and all(len(g) == 2 * N * 8 for g in got)
    xpad = np.concatenate([x0, np.zeros(2 * N - len(x0))])
    for c, xin in enumerate([xpad, np.zeros(2 * N)]):
        ref = np.concatenate([ochains[c].process(xin[k * N:(k + 1) * N], fs) for k in range(2)])
        y = np.frombuffer(got[c], dtype=np.float64)
        for k in range(2):
            assert common.rms(ref[k * N:(k + 1) * N] - y[k * N:(k + 1) * N]) <= common.TOL_RMS, (c, k)
    assert common.rms(np.frombuffer(got[0], dtype=np.float64)) > 1e-4
    assert got[2] == got[3] == got[4] == bytes(2 * N * 8)
