"""BASELINE.json's headline configuration at FULL size on the GPU (cfg5: 1024 channels, CHAIN12 with 4x oversampling and an
8192-tap power-amp filter, 192 kHz, blocks of 8192, spatializer on, meter on): the oracle is far too slow to follow all
1024 channels, so the batch is checked through
  * direct parity on a spread of channels (first / last channel, both sides of every SM-wave boundary of the pipeline),
  * properties that do not depend on the size: channels that are fed the same input produce the same bits, a channel's
    output does not depend on the batch it travels in, the stereo master is the oracle's spatializer applied to the
    1024 outputs, the meters agree with the oracle's on the checked channels, two blocks in flight change nothing."""
import numpy as np
import pytest

import common
from common import gdg, orc

pytestmark = pytest.mark.gpu

FS, N, C, BLOCKS, PATTERNS = 192000, 8192, 1024, 2, 64
CHECK = [0, 1, 63, 64, 147, 148, 149, 295, 296, 511, 512, 739, 740, 887, 888, 1023]


def _inputs():
    base = np.stack([common.synth_fast(p, FS, BLOCKS * N, 3) for p in range(PATTERNS)])
    return [np.ascontiguousarray(base[np.arange(C) % PATTERNS, k * N:(k + 1) * N]) for k in range(BLOCKS)]


def _chains(lib, bank, count):
    return [common.build_chain(gdg.Chain(bank, lib=lib), common.CHAIN12, "syn8192") for _ in range(count)]


def test_cfg5_full_batch():
    lib = common.backend_lib("cuda")
    ob, pb = common.make_banks(lib, with_fixture=False, synthetic=[("syn8192", 8192, FS)])
    x = _inputs()
    chains = _chains(lib, pb, C)
    spat, ospat = gdg.Spatializer(C, lib=lib), orc.Spatializer(C)
    for c in range(C):
        az, dist, lv = -90.0 + 180.0 * c / (C - 1), 1.0 + (c % 7) * 0.5, 1.0 / (1 + c % 5)
        for s in (spat, ospat):
            s.set_azimuth(c, az); s.set_distance(c, dist); s.set_level(c, lv)
    meter, ometer = gdg.Meter(2 * C + 3, lib=lib), orc.Meter(2 * len(CHECK))
    meter.set_enabled(True); ometer.set_enabled(True)

    # the full batch, two blocks in flight
    outs = [np.zeros((C + 2, N)) for _ in range(BLOCKS)]
    tickets = [gdg.controller_submit(chains, None, -1, spat, x[k], outs[k], FS, meter=meter) for k in range(BLOCKS)]
    for t in tickets:
        gdg.controller_wait(t, lib)
    for k in range(BLOCKS):
        assert np.all(np.isfinite(outs[k]))

    # (1) parity on the spread of channels + their meters
    ochains = [common.build_chain(orc.Chain(ob), common.CHAIN12, "syn8192") for _ in CHECK]
    for k in range(BLOCKS):
        ref = [oc.process(x[k][c], FS) for oc, c in zip(ochains, CHECK)]
        for r, c in zip(ref, CHECK):
            assert common.rms(r - outs[k][c]) <= common.TOL_RMS, (k, c)
        ometer.process([x[k][c] for c in CHECK] + ref, FS)
    for i, c in enumerate(CHECK):
        assert meter.analyze(c) == ometer.analyze(i) and meter.analyze(C + c) == ometer.analyze(len(CHECK) + i), c
        assert meter.state(c)[2] == ometer.state(i)[2]

    # (2) equal inputs -> equal bits, across CTAs, waves and pipeline groups
    for k in range(BLOCKS):
        y = outs[k][:C].reshape(C // PATTERNS, PATTERNS, N)
        assert np.array_equal(y, np.broadcast_to(y[0], y.shape)), k

    # (3) the master pair is the spatializer over all 1024 outputs (group partials: order differs, 1e-12 relative)
    for k in range(BLOCKS):
        m = ospat.process(outs[k][:C], None)
        scale = max(1.0, float(np.max(np.abs(m))))
        assert np.max(np.abs(m - outs[k][C:])) <= 1e-11 * scale, k

    # (4) a channel does not care about the batch it travels in, nor about blocking vs. in-flight submission
    small = _chains(lib, pb, len(CHECK))
    for k in range(BLOCKS):
        ys = gdg.process_batch(small, x[k][CHECK], FS)
        assert np.array_equal(ys, outs[k][CHECK]), k


def test_cfg5_device_resident_batch_on_two_kernel_streams():
    """gdg_process_batch_device at full size: the two halves of the channels run on two kernel streams inside the call
    (forked from and joined to the caller's stream).  Every channel must produce the bits it produces in a small batch on one
    stream, block after block, and the caller's stream must still order consecutive calls (block k + 1 reads the state block k
    wrote)."""
    import torch
    lib = common.backend_lib("cuda")
    ob, pb = common.make_banks(lib, with_fixture=False, synthetic=[("syn8192", 8192, FS)])
    x = _inputs()
    chains = _chains(lib, pb, C)
    small = _chains(lib, pb, len(CHECK))
    stream = torch.cuda.Stream()
    runner = gdg.BatchRunner(chains)
    xin = [torch.tensor(x[k], device="cuda") for k in range(BLOCKS)]
    yout = [torch.empty_like(xin[k]) for k in range(BLOCKS)]
    torch.cuda.synchronize()
    for k in range(BLOCKS):                                   # both blocks enqueued back to back, no host synchronisation in between
        runner.process_device(xin[k].data_ptr(), yout[k].data_ptr(), N, N, FS, stream.cuda_stream)
    stream.synchronize()
    for k in range(BLOCKS):
        y = yout[k].cpu().numpy()
        ys = gdg.process_batch(small, x[k][CHECK], FS)
        assert np.array_equal(ys, y[CHECK]), k
        yy = y.reshape(C // PATTERNS, PATTERNS, N)
        assert np.array_equal(yy, np.broadcast_to(yy[0], yy.shape)), k
