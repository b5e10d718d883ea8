#!/usr/bin/env python3
"""Run one chain spec a few times (for ncu): tools/one_unit.py <unit|CHAIN12> [C]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import common
from common import gdg
lib = common.backend_lib("cuda")
unit = sys.argv[1]; C = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
n, fs = 8192, 192000
banks = common.make_banks(lib, with_fixture=False, synthetic=[("syn8192", 8192, fs)])
spec = common.CHAIN12 if unit == "CHAIN12" else [s for s in common.CHAIN12 if s[0] == unit] or [(unit, {}, {})]
chains = [common.build_chain(gdg.Chain(banks[1], lib=lib), spec, "syn8192") for _ in range(C)]
r = gdg.BatchRunner(chains)
x = torch.tensor(np.stack([common.synth_fast(c % 16, fs, n) for c in range(16)]), device="cuda").repeat(C // 16 + 1, 1)[:C].contiguous()
y = torch.empty_like(x)
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
for _ in range(4):
    r.process_device(x.data_ptr(), y.data_ptr(), n, n, fs, ts.cuda_stream)
torch.cuda.synchronize()
print("done", unit)
