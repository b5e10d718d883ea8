"""World-size-2 CPU test (gloo) of the multi-GPU path's host logic: channel sharding and the
stereo-master reduce.  Each rank runs its shard on the host-emulated backend; rank 0 checks the
reduced master and the gathered channel outputs against the single-process oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import common

C_TOTAL, N, FS, BLOCKS = 6, 1024, 96000, 3
SPEC = [("overdrive", {"gain": 15}, {"oversampling": "2"}), ("tone_stack", {}, {}), ("delay", {"delay_time": 5}, {})]


def _positions(c):
    return -60.0 + 25.0 * c, 0.5 + 0.3 * c, 1.0 - 0.1 * c


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib
    par = importlib.import_module("go-dsp-guitar_b200.parallel")
    from common import gdg
    lib = common.backend_lib("emu")
    lo, hi = par.shard_channels(C_TOTAL, world, rank)
    chains = [common.build_chain(gdg.Chain(lib=lib), SPEC) for _ in range(lo, hi)]
    spat = gdg.Spatializer(hi - lo, lib=lib)
    spat.set_sample_rate(FS)
    for c in range(lo, hi):
        az, d, lv = _positions(c)
        spat.set_azimuth(c - lo, az); spat.set_distance(c - lo, d); spat.set_level(c - lo, lv)
    outs = []
    for k in range(BLOCKS):
        x = np.stack([common.synth_fast(c, FS, BLOCKS * N)[k * N:(k + 1) * N] for c in range(lo, hi)])
        y = gdg.controller_process(chains, None, -1, spat, x, FS)
        master = torch.from_numpy(y[hi - lo:].copy())
        par.reduce_stereo_master(master, dst=0)
        gathered = [None] * world
        dist.all_gather_object(gathered, y[:hi - lo])
        if rank == 0:
            outs.append(np.concatenate(gathered + [master.numpy()]))
    if rank == 0:
        ret.put(np.stack(outs))
    dist.destroy_process_group()


def test_two_rank_sharding_and_master_reduce():
    from common import orc
    import importlib
    par = importlib.import_module("go-dsp-guitar_b200.parallel")
    assert [par.shard_channels(7, 3, r) for r in range(3)] == [(0, 3), (3, 5), (5, 7)]
    assert par.owner_of_channel(4, 7, 3) == 1
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    [p.start() for p in procs]
    got = ret.get(timeout=120)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    # single-process oracle over all channels
    chains = [common.build_chain(orc.Chain(), SPEC) for _ in range(C_TOTAL)]
    spat = orc.Spatializer(C_TOTAL)
    spat.set_sample_rate(FS)
    for c in range(C_TOTAL):
        az, d, lv = _positions(c)
        spat.set_azimuth(c, az); spat.set_distance(c, d); spat.set_level(c, lv)
    for k in range(BLOCKS):
        x = np.stack([common.synth_fast(c, FS, BLOCKS * N)[k * N:(k + 1) * N] for c in range(C_TOTAL)])
        ref = orc.controller_process(chains, None, -1, spat, x, FS)
        assert got[k].shape == ref.shape
        for c in range(C_TOTAL + 2):
            assert common.rms(got[k][c] - ref[c]) <= common.TOL_RMS, (k, c)
