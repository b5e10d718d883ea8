#!/usr/bin/env python3
"""Per-unit and CHAIN12 device-resident timings (development aid; bench.py is the contract)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import common
from common import gdg

lib = common.backend_lib("cuda")
C = int(os.environ.get("GDG_C", "1024")); n = int(os.environ.get("GDG_N", "8192")); fs = 192000
banks = common.make_banks(lib, synthetic=[("syn8192", 8192, fs)])
x = torch.tensor(np.stack([common.synth_fast(c % 64, fs, n) for c in range(64)]), device="cuda").repeat(C // 64 + 1, 1)[:C].contiguous()
y = torch.empty_like(x)
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
stream = ts.cuda_stream
assert stream
res = {}
def timeit(name, spec, ir=None, reps=int(os.environ.get("GDG_REPS", "5"))):
    chains = [common.build_chain(gdg.Chain(banks[1], lib=lib), spec, ir) for _ in range(C)]
    r = gdg.BatchRunner(chains)
    for _ in range(2): r.process_device(x.data_ptr(), y.data_ptr(), n, n, fs, stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r.process_device(x.data_ptr(), y.data_ptr(), n, n, fs, stream)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    res[name] = ms
    print("%-28s %8.3f ms/block  %9.1f Msamples/s" % (name, ms, C * n / ms / 1e3), flush=True)
only = os.environ.get("GDG_ONLY", "").split(",") if os.environ.get("GDG_ONLY") else None
timeit("empty", [])
for u, nn, dd in common.CHAIN12:
    if only and u not in only: continue
    timeit(u, [(u, nn, dd)], "syn8192")
for u in ["signal_generator", "bandpass", "auto_wah", "auto_yoy", "octaver", "flanger", "phaser", "tremolo", "ring_modulator"]:
    if only and u not in only: continue
    timeit(u, [(u, {}, {})])
timeit("CHAIN12", common.CHAIN12, "syn8192")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"C": C, "n": n, "fs": fs, "ms": res}, open(os.path.join(ROOT, "gpurun_out", "quick_bench.json"), "w"), indent=1)
