"""Several GPUs in ONE process (gdg_cluster_*): channel ranges per device, tuner with its channel's device, the stereo
master summed on member 0 from the other members' device buffers (peer loads on real hardware).

CPU: the emulation exposes GDG_EMU_DEVICES fake devices (one memory space) - this checks the host logic: context
selection per object, channel partition, per-member pipelines, tickets.  GPU: the same test on two real B200s (skipped on
one-GPU boxes; run with `gpurun --gpus 2`)."""
import os

import numpy as np
import pytest

os.environ.setdefault("GDG_EMU_DEVICES", "3")
import common  # noqa: E402
from common import gdg, orc  # noqa: E402

BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]
SPEC = [("compressor", {}, {}), ("overdrive", {"gain": 14}, {"oversampling": "4"}), ("chorus", {}, {}), ("delay", {"delay_time": 7}, {}),
        ("power_amp", {"level_1": -4}, {"filter_order": "512", "filter_1": "ir"}), ("cabinet", {}, {})]


@pytest.fixture(params=BACKENDS)
def lib(request):
    return common.backend_lib(request.param)


def _build(lib, devices, counts, fs):
    cl = gdg.Cluster(devices, lib=lib)
    ir = common.synthetic_ir(700)
    chains, spats, ochains = [], [], []
    ob = orc.IRBank()
    ob.add_raw("ir", -2, ir, fs)
    g = 0
    total = sum(counts)
    ospat = orc.Spatializer(total)
    ospat.set_sample_rate(fs)
    for m, cnt in enumerate(counts):
        cl.select(m)
        bank = gdg.IRBank(lib)                       # responses live on the member's device
        bank.add_raw("ir", -2, ir, fs)
        sp = gdg.Spatializer(cnt, lib=lib)
        sp.set_sample_rate(fs)
        for c in range(cnt):
            chains.append(common.build_chain(gdg.Chain(bank, lib=lib), SPEC, "ir"))
            ochains.append(common.build_chain(orc.Chain(ob), SPEC, "ir"))
            az, dist = -80.0 + 160.0 * g / max(1, total - 1), 0.8 + 0.3 * (g % 5)
            sp.set_azimuth(c, az); sp.set_distance(c, dist)
            ospat.set_azimuth(g, az); ospat.set_distance(g, dist)
            g += 1
        spats.append(sp)
        chains[-1]._keep = bank
    return cl, chains, spats, ochains, ospat


def test_cluster_matches_single_stream_reference(lib):
    ndev = lib.gdg_device_count()
    if ndev < 2:
        pytest.skip("needs two devices (gpurun --gpus 2)")
    devices = list(range(min(ndev, 3)))
    counts = [3, 2, 4][:len(devices)]
    fs, n, blocks = 96000, 1024, 5
    cl, chains, spats, ochains, ospat = _build(lib, devices, counts, fs)
    C = sum(counts)
    cl.select(len(devices) - 1)
    tuner = gdg.Tuner(lib=lib)                       # the tuner lives with its channel's device (the last member here)
    otuner = orc.Tuner()
    tch = C - 1
    x = np.stack([common.synth_fast(c, fs, blocks * n, 2) for c in range(C)])
    for k in range(blocks):
        blk = np.ascontiguousarray(x[:, k * n:(k + 1) * n])
        got = cl.process(chains, tuner, tch, spats, blk, fs)
        ref = orc.controller_process(ochains, otuner, tch, ospat, blk, fs)
        for c in range(C):
            assert common.rms(got[c] - ref[c]) <= common.TOL_RMS, (k, c)
        # the master is summed per member, then over members: the order differs from the reference's 0..C-1 (SURVEY 8e caveat)
        assert np.max(np.abs(got[C:] - ref[C:])) <= 1e-12 * max(1.0, np.max(np.abs(ref[C:]))), k
        assert np.max(np.abs(ref[C:])) > 0
    assert lib.gdg_current_device() == devices[-1]   # cluster calls restore the caller's selection
    lib.gdg_use_device(devices[0])


def test_cluster_rejects_misplaced_chains(lib):
    ndev = lib.gdg_device_count()
    if ndev < 2:
        pytest.skip("needs two devices")
    cl = gdg.Cluster([0, 1], lib=lib)
    cl.select(1)
    a = common.build_chain(gdg.Chain(lib=lib), [("distortion", {}, {})])
    cl.select(0)
    b = common.build_chain(gdg.Chain(lib=lib), [("distortion", {}, {})])
    x = np.zeros((2, 256))
    with pytest.raises(gdg.GdgError, match="member order"):
        cl.process([a, b], None, -1, None, x, 48000)               # device 1's chain listed before device 0's
    with pytest.raises(gdg.GdgError, match="one device"):
        gdg.process_batch([a, b], x, 48000)                        # a plain batch cannot span devices
    assert np.array_equal(cl.process([b, a], None, -1, None, x + 0.25, 48000)[0], np.full(256, 0.25))
    lib.gdg_use_device(0)
