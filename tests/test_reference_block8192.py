"""The headline configuration against the EXECUTED reference: CHAIN12 at 192 kHz in the batch path's own block length of 8192
(controller.go:36), power amps on "Guitar: Custom BP" (filter.Reduce 15359 -> 8192 taps) and "Guitar: American Vintage (Center)",
8 consecutive blocks per channel, then the reference's spatializer - recorded from the reference's Go sources run through go2py
(tests/golden/go2py/make_block8192_vectors.py, needs the reference tree; the vectors do not).  The other recorded streams use
blocks of 32 ... 2048; this one pins the block length, rate and chain BASELINE.json's metric is quoted on."""
import os

import numpy as np
import pytest

import common
from common import gdg, orc

Z = np.load(os.path.join(common.GOLD, "translated_reference_block8192.npz"))
FS, N, BLOCKS = int(Z["fs"]), int(Z["n"]), int(Z["blocks"])
IRS = [str(s) for s in Z["irs"]]
GEOMETRY = [tuple(float(v) for v in g) for g in Z["geometry"]]
REF_OUT, REF_MASTER = Z["outputs"], Z["master"]


def _inputs():
    return np.stack([common.synth_channel(c, FS, N * BLOCKS) for c in range(len(IRS))])      # SURVEY.md 8d, regenerated


def test_oracle_reproduces_the_executed_reference_at_the_batch_block_length():
    ob, _ = common.make_banks(common.backend_lib("emu"))
    chains = [common.build_chain(orc.Chain(ob), common.CHAIN12, ir) for ir in IRS]
    sp = orc.Spatializer(len(IRS))
    sp.set_sample_rate(FS)
    for c, (az, dist, lvl) in enumerate(GEOMETRY):
        sp.set_azimuth(c, az); sp.set_distance(c, dist); sp.set_level(c, lvl)
    x = _inputs()
    assert np.abs(REF_OUT).max() > 5e-3 and np.abs(REF_MASTER).max() > 5e-3                  # the streams carry signal
    for k in range(BLOCKS):
        sl = slice(k * N, (k + 1) * N)
        y = np.stack([chains[c].process(x[c, sl], FS) for c in range(len(IRS))])
        assert np.max(np.abs(y - REF_OUT[:, sl])) <= 1e-12 * max(1.0, np.abs(REF_OUT).max()), k
        assert np.max(np.abs(sp.process(y) - REF_MASTER[:, sl])) <= 1e-12, k


@pytest.mark.parametrize("backend", [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)])
def test_product_reproduces_the_executed_reference_at_the_batch_block_length(backend):
    lib = common.backend_lib(backend)
    _, pb = common.make_banks(lib)
    chains = [common.build_chain(gdg.Chain(pb, lib=lib), common.CHAIN12, ir) for ir in IRS]
    sp = gdg.Spatializer(len(IRS), lib=lib)
    sp.set_sample_rate(FS)
    for c, (az, dist, lvl) in enumerate(GEOMETRY):
        sp.set_azimuth(c, az); sp.set_distance(c, dist); sp.set_level(c, lvl)
    x = _inputs()
    C = len(IRS)
    for k in range(BLOCKS):
        sl = slice(k * N, (k + 1) * N)
        rows = gdg.controller_process(chains, None, -1, sp, x[:, sl], FS)                    # one call: chains + spatializer (controller.go:2648-2783)
        for c in range(C):
            e, r = common.rms(rows[c] - REF_OUT[c, sl]), common.rms(REF_OUT[c, sl])
            assert e <= common.TOL_RMS and e <= 1e-8 * r and r > 0, (k, c, e, r)
        for j in range(2):
            e, r = common.rms(rows[C + j] - REF_MASTER[j, sl]), common.rms(REF_MASTER[j, sl])
            assert e <= common.TOL_RMS and e <= 1e-8 * r and r > 0, (k, "master", j, e, r)
