#!/usr/bin/env python3
"""Control-path timing: a power-amp parameter change on 1024 chains, compiled on the device.
  same     : every chain gets the same new level -> ONE composite filter + spectrum, 1023 cache hits
  distinct : every chain gets its own level / IR mix -> 1024 compilations (sum of squares, scale-add, spectra; Reduce for
             the responses longer than the target order)
Prints one JSON object (and writes gpurun_out/recompile_bench.json)."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
gdg = importlib.import_module("go-dsp-guitar_b200")
presets = importlib.import_module("go-dsp-guitar_b200.presets")
lib = gdg.load(); gdg.init(-1, lib)
C, fs, n = 1024, 192000, 8192
bank = gdg.IRBank(lib)
t0 = time.perf_counter()
for k in range(24):                                   # 24 responses x 7 rates, like ir/index.json (filter.go:704-802)
    bank.add("ir%02d" % k, -k % 5, presets.synthetic_ir(511 + 380 * k, seed=k + 1), 96000)
lib.gdg_synchronize()
t_import = time.perf_counter() - t0
chains = []
for c in range(C):
    ch = gdg.Chain(bank, lib=lib)
    i = ch.append_unit("power_amp")
    ch.set_discrete_value(i, "filter_order", "8192")
    ch.set_discrete_value(i, "filter_1", "ir23")      # 9251 frames @96k -> 18501 taps @192k: reduced to 8192
    ch.set_discrete_value(i, "filter_2", "ir05")
    ch.set_bypass(i, False)
    chains.append(ch)
import torch
x = torch.zeros((C, n), dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
runner = gdg.BatchRunner(chains)
def run_block():
    runner.process_device(x.data_ptr(), y.data_ptr(), n, n, fs, None); lib.gdg_synchronize()
t0 = time.perf_counter(); run_block(); t_first = time.perf_counter() - t0          # first block: rate known -> compile
def retime(label, setter):
    t0 = time.perf_counter()
    for c, ch in enumerate(chains): setter(c, ch)
    t_set = time.perf_counter() - t0
    t0 = time.perf_counter(); run_block(); t_blk = time.perf_counter() - t0
    return {"set_calls_ms": 1e3 * t_set, "next_block_ms": 1e3 * t_blk, "total_ms": 1e3 * (t_set + t_blk)}
out = {"channels": C, "import_24_responses_x7_rates_ms": 1e3 * t_import, "first_block_ms": 1e3 * t_first,
       "same_level_on_all_chains": retime("same", lambda c, ch: ch.set_numeric_value(0, "level_1", -6)),
       "distinct_levels": retime("distinct", lambda c, ch: (ch.set_numeric_value(0, "level_1", -(c % 60)), ch.set_numeric_value(0, "level_2", -((c // 60) % 60)))),
       "steady_block_ms": None}
t0 = time.perf_counter(); run_block(); out["steady_block_ms"] = 1e3 * (time.perf_counter() - t0)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "recompile_bench.json"), "w"), indent=1)
print(json.dumps(out))
