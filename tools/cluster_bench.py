#!/usr/bin/env python3
"""ONE process driving N GPUs through gdg_cluster_* (no torchrun): e2e Msamples/s and the host time inside submit / wait.
usage: tools/cluster_bench.py [ndev] [channels] [blocks]"""
import ctypes as Ct, importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
gdg = importlib.import_module("go-dsp-guitar_b200"); presets = importlib.import_module("go-dsp-guitar_b200.presets")
lib = gdg.load(); gdg.init(0, lib)
ndev = int(sys.argv[1]) if len(sys.argv) > 1 else lib.gdg_device_count()
C = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
blocks = int(sys.argv[3]) if len(sys.argv) > 3 else 128
FS, N = 192000, 8192
cl = gdg.Cluster(list(range(ndev)), lib=lib)
per, rem = divmod(C, ndev)
chains, spats, keep = [], [], []
g0 = 0
for m in range(ndev):
    cnt = per + (1 if m < rem else 0)
    cl.select(m)
    bank = gdg.IRBank(lib); bank.add_raw("syn8192", 0, presets.synthetic_ir(8192), FS); keep.append(bank)
    sp = gdg.Spatializer(cnt, lib=lib); sp.set_sample_rate(FS)
    for c in range(cnt):
        chains.append(presets.build_chain(gdg.Chain(bank, lib=lib), presets.CHAIN12, "syn8192"))
        sp.set_azimuth(c, -70.0 + 140.0 * (g0 + c) / max(1, C - 1)); sp.set_distance(c, 1.0 + 0.25 * ((g0 + c) % 8))
    spats.append(sp); g0 += cnt
cl.select(0)
tuner = gdg.Tuner(lib=lib)
h_in = [torch.empty((C, N), dtype=torch.float64).pin_memory() for _ in range(2)]
h_out = [torch.empty((C + 2, N), dtype=torch.float64).pin_memory() for _ in range(2)]
base = torch.tensor(np.stack([presets.synth_fast(c, FS, N) for c in range(36)]))
for k in range(2): h_in[k].copy_(base[torch.arange(C) % 36])
handles = (Ct.c_void_p * C)(*[c_._h for c_ in chains]); sph = (Ct.c_void_p * ndev)(*[s_._h for s_ in spats])
pin = [(Ct.c_void_p * C)(*[h_in[k].data_ptr() + 8 * N * c for c in range(C)]) for k in range(2)]
pout = [(Ct.c_void_p * (C + 2))(*[h_out[k].data_ptr() + 8 * N * c for c in range(C + 2)]) for k in range(2)]
def run(nb):
    ts = tw = 0.0; prev = None
    for k in range(nb):
        a = time.perf_counter()
        t = lib.gdg_cluster_submit(cl._h, handles, C, tuner._h, 0, sph, pin[k % 2], pout[k % 2], C + 2, N, FS)
        assert t >= 0, lib.gdg_last_error()
        b = time.perf_counter(); ts += b - a
        if prev is not None:
            assert lib.gdg_cluster_wait(cl._h, prev) >= 0
            tw += time.perf_counter() - b
        prev = t
    assert lib.gdg_cluster_wait(cl._h, prev) >= 0
    return ts, tw
run(8)
t0 = time.perf_counter(); ts, tw = run(blocks); dt = time.perf_counter() - t0
print({"ndev": ndev, "channels": C, "blocks": blocks, "msamples_per_s": C * N * blocks / dt / 1e6, "ms_per_block": 1e3 * dt / blocks,
       "host_ms_in_submit_per_block": 1e3 * ts / blocks, "host_ms_in_wait_per_block": 1e3 * tw / blocks})
