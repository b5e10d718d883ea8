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


METER_RTOL = 1e-10


def _cluster_meters(lib, cl, counts):
    """one level meter per member: its inputs, its outputs, metronome port, master L / R (include/gdg.h)"""
    meters = []
    for m, cnt in enumerate(counts):
        cl.select(m)
        mt = gdg.Meter(2 * cnt + 3, lib=lib)
        mt.set_enabled(True)
        meters.append(mt)
    return meters


def _assert_cluster_meters_agree(meters, counts, ometer):
    """the reference's ONE meter of 2C + 3 channels (controller.go:2710-2771) against the members' meters"""
    C = sum(counts)
    first = np.concatenate([[0], np.cumsum(counts)])
    pairs = []
    for m, cnt in enumerate(counts):
        for c in range(cnt):
            pairs.append((meters[m], c, int(first[m]) + c))                     # inputs
            pairs.append((meters[m], cnt + c, C + int(first[m]) + c))           # chain outputs
    for k in range(3):
        pairs.append((meters[0], 2 * counts[0] + k, 2 * C + k))                 # metronome, master L / R: member 0
    for mt, lc, gc in pairs:
        (mc, mp, mq), (oc, op, oq) = mt.state(lc), ometer.state(gc)
        assert mq == oq, (gc, mq, oq)
        assert abs(mc - oc) <= METER_RTOL * abs(oc) + 1e-300 and abs(mp - op) <= METER_RTOL * abs(op) + 1e-300, (gc, mc, oc, mp, op)
        assert mt.analyze(lc) == ometer.analyze(gc)
    for m in range(1, len(counts)):                                             # the other members never feed their last three
        for k in range(3):
            assert meters[m].state(2 * counts[m] + k)[:2] == (0.0, 0.0)


def test_cluster_level_meters_and_metronome(lib):
    """controller.go:2710-2771 over several devices: per-member meters, the metronome row into meter port 2C and - with
    metrMasterOutput - into the master inside the peer-memory sum (spatializer.go:300-308)."""
    ndev = lib.gdg_device_count()
    if ndev < 2:
        pytest.skip("needs two devices (gpurun --gpus 2)")
    devices = list(range(min(ndev, 3)))
    counts = [3, 2, 4][:len(devices)]
    fs, n, blocks = 48000, 512, 6
    cl, chains, spats, ochains, ospat = _build(lib, devices, counts, fs)
    C = sum(counts)
    meters = _cluster_meters(lib, cl, counts)
    ometer = orc.Meter(2 * C + 3)
    ometer.set_enabled(True)
    rng = np.random.default_rng(11)
    x = np.stack([common.synth_fast(c, fs, blocks * n, 2) for c in range(C)])
    for k in range(blocks):
        blk = np.ascontiguousarray(x[:, k * n:(k + 1) * n])
        click = np.where(np.arange(n) < 48, 0.3 * rng.uniform(-1, 1, n), 0.0)
        to_master = k % 2 == 0
        got = cl.process(chains, None, -1, spats, blk, fs, meters=meters, metronome=click, metronome_to_master=to_master)
        outs = np.stack([ochains[c].process(blk[c], fs) for c in range(C)])
        master = ospat.process(outs, click if to_master else None)
        assert np.max(np.abs(got[:C] - outs)) <= 1e-9
        assert np.max(np.abs(got[C:] - master)) <= 1e-12 * max(1.0, np.max(np.abs(master))), k
        ometer.process(list(blk) + list(outs) + [click] + list(master), fs)
    _assert_cluster_meters_agree(meters, counts, ometer)
    assert meters[0].analyze(2 * counts[0])[0] > -60         # the metronome port saw the click
    with pytest.raises(gdg.GdgError, match="level meter must have"):
        cl.process(chains, None, -1, spats, blk, fs, meters=[meters[1], meters[0]] + meters[2:])
    lib.gdg_use_device(devices[0])


def test_cluster_two_blocks_in_flight_keep_their_masters(lib):
    """submit(k), submit(k+1), wait(k), wait(k+1): each ticket's master pair is its own block's (per-ticket master buffers)."""
    ndev = lib.gdg_device_count()
    if ndev < 2:
        pytest.skip("needs two devices (gpurun --gpus 2)")
    devices = list(range(min(ndev, 3)))
    counts = [2, 3, 1][:len(devices)]
    fs, n, blocks = 96000, 256, 6
    cl, chains, spats, ochains, ospat = _build(lib, devices, counts, fs)
    C = sum(counts)
    x = np.stack([common.synth_fast(c, fs, blocks * n, 2) for c in range(C)])
    ins = [np.ascontiguousarray(x[:, k * n:(k + 1) * n]) for k in range(blocks)]
    outs = [np.zeros((C + 2, n)) for _ in range(blocks)]
    refs = [orc.controller_process(ochains, None, -1, ospat, ins[k], fs) for k in range(blocks)]
    prev = None
    for k in range(blocks):
        t = cl.submit(chains, None, -1, spats, ins[k], outs[k], fs)
        if prev is not None:
            cl.wait(prev)
        prev = t
    cl.wait(prev)
    with pytest.raises(gdg.GdgError, match="not in flight"):
        cl.wait(prev)
    for k in range(blocks):
        assert np.max(np.abs(outs[k][:C] - refs[k][:C])) <= 1e-9, k
        assert np.max(np.abs(outs[k][C:] - refs[k][C:])) <= 1e-12 * max(1.0, np.max(np.abs(refs[k][C:]))), k
        assert np.max(np.abs(refs[k][C:])) > 0
    lib.gdg_use_device(devices[0])


@pytest.mark.parametrize("fmt", ["PCM16", "PCM24", "IEEE32"])
def test_cluster_encoded_rows(lib, fmt):
    """controller.go:2908-3143 for one block over several devices without a float64 crossing PCIe: every member decodes /
    encodes its rows in its segment kernels; the master pair is encoded on member 0 (its summation order differs from 0..C-1 by
    <= 1 ulp of the sum: compared after decoding)."""
    ndev = lib.gdg_device_count()
    if ndev < 2:
        pytest.skip("needs two devices (gpurun --gpus 2)")
    devices = list(range(min(ndev, 3)))
    counts = [2, 1, 3][:len(devices)]
    fs, n, blocks = 48000, 1024, 4
    cl, chains, spats, ochains, ospat = _build(lib, devices, counts, fs)
    C = sum(counts)
    meters = _cluster_meters(lib, cl, counts)
    ometer = orc.Meter(2 * C + 3)
    ometer.set_enabled(True)
    bps = lib.gdg_format_bytes_per_sample(gdg.FORMATS.index(fmt))
    x = np.stack([common.synth_fast(c, fs, blocks * n, 2) for c in range(C)])
    for k in range(blocks):
        blk = x[:, k * n:(k + 1) * n]
        rows = [np.frombuffer(orc.wave_encode(fmt, blk[c]).tobytes(), dtype=np.uint8).copy() for c in range(C)]
        outs = [np.zeros(n * bps, dtype=np.uint8) for _ in range(C + 2)]
        cl.wait(cl.submit_encoded(chains, None, -1, spats, rows, fmt, outs, fmt, n, fs, meters=meters))
        dec = np.stack([orc.wave_decode(fmt, rows[c].tobytes()) for c in range(C)])
        y = np.stack([ochains[c].process(dec[c], fs) for c in range(C)])
        master = ospat.process(y, None)
        step = {"PCM16": 2.0 ** -15, "PCM24": 2.0 ** -23, "IEEE32": 2.0 ** -23}[fmt]
        for c in range(C):                     # the oracle's own encoder on the oracle's float64 result; the product's float64
            want = orc.wave_encode(fmt, y[c])  # differs by ~1e-16 absolute, which may move a sample across one rounding boundary
            if outs[c].tobytes() != want.tobytes():
                d = np.abs(orc.wave_decode(fmt, outs[c].tobytes()) - orc.wave_decode(fmt, want.tobytes()))
                if fmt == "IEEE32":
                    assert np.max(d) <= 1e-14 + 2.0 ** -23 * np.max(np.abs(y[c])) and np.count_nonzero(d) <= n // 16, (k, c)
                else:
                    assert np.max(d) <= 1.001 * step and np.count_nonzero(d) <= 2, (k, c)
        gm = np.stack([orc.wave_decode(fmt, outs[C + r].tobytes()) for r in range(2)])
        assert np.max(np.abs(gm - np.clip(master, -1.0, 1.0))) <= 1.001 * step * max(1.0, np.max(np.abs(master))), k
        assert np.max(np.abs(gm)) > 0
        ometer.process(list(dec) + list(y) + [np.zeros(n)] + list(master), fs)
    _assert_cluster_meters_agree(meters, counts, ometer)
    lib.gdg_use_device(devices[0])


def test_cluster_block_length_grows_while_a_block_is_in_flight(lib):
    """submit(n = 128), submit(n = 512), then the waits: the first ticket's master rows must survive the second submit (the
    cluster's master / staging buffers are sized once for the longest block, never re-made under a ticket in flight)"""
    ndev = lib.gdg_device_count()
    if ndev < 2:
        pytest.skip("needs two devices (gpurun --gpus 2)")
    devices = list(range(min(ndev, 3)))
    counts = [2, 2, 1][:len(devices)]
    fs = 48000
    cl, chains, spats, ochains, ospat = _build(lib, devices, counts, fs)
    C = sum(counts)
    lens = [128, 512, 256, 1024]
    x = np.stack([common.synth_fast(c, fs, sum(lens), 2) for c in range(C)])
    ins, outs, refs, pos = [], [], [], 0
    for n in lens:
        ins.append(np.ascontiguousarray(x[:, pos:pos + n])); outs.append(np.zeros((C + 2, n))); pos += n
        refs.append(orc.controller_process(ochains, None, -1, ospat, ins[-1], fs))
    prev = None
    for k in range(len(lens)):
        t = cl.submit(chains, None, -1, spats, ins[k], outs[k], fs)
        if prev is not None:
            cl.wait(prev)
        prev = t
    cl.wait(prev)
    for k in range(len(lens)):
        assert np.max(np.abs(outs[k][:C] - refs[k][:C])) <= 1e-9, k
        assert np.max(np.abs(outs[k][C:] - refs[k][C:])) <= 1e-12 * max(1.0, np.max(np.abs(refs[k][C:]))), k
    lib.gdg_use_device(devices[0])
