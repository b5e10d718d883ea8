#!/usr/bin/env python3
"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): every unit type, CHAIN12 with a partitioned
power-amp filter, spatializer, tuner, meter, codecs - a few channels, two blocks each.
    compute-sanitizer --tool racecheck python tools/sanitize.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import common
from common import gdg
lib = common.backend_lib("cuda")
fs, n, C = 192000, 8192, 3
ob, pb = common.make_banks(lib, with_fixture=False, synthetic=[("syn8192", 8192, fs), ("syn20000", 20000, fs)])
every = [(u, {}, ({"oversampling": "2"} if u in ("excess", "fuzz") else {})) for u in gdg.UNIT_TYPES]
every[19] = ("power_amp", {}, {"filter_order": "32768", "filter_1": "syn20000"})
specs = [common.CHAIN12, every, [("fuzz", {}, {"oversampling": "4"}), ("bandpass", {}, {"filter_order": "8"}), ("octaver", {}, {}), ("auto_wah", {}, {})]]
chains = [common.build_chain(gdg.Chain(pb, lib=lib), specs[c], "syn8192") for c in range(C)]
spat, tuner, meter = gdg.Spatializer(C, lib=lib), gdg.Tuner(lib=lib), gdg.Meter(2 * C + 3, lib=lib)
meter.set_enabled(True)
for k in range(2):
    x = np.stack([common.synth_fast(c, fs, 2 * n)[k * n:(k + 1) * n] for c in range(C)])
    y = gdg.controller_process(chains, tuner, 0, spat, x, fs, meter=meter)
    assert np.all(np.isfinite(y))
small = [common.build_chain(gdg.Chain(pb, lib=lib), [("overdrive", {}, {"oversampling": "4"}), ("reverb", {}, {})]) for _ in range(2)]
for nn in (64, 1024):
    rows = [(np.sin(np.arange(nn) * 0.05 * (c + 1)) * 20000).astype("<i2").tobytes() for c in range(2)]
    gdg.process_batch_encoded(small, rows, "PCM16", "PCM24", nn, 44100)
print(tuner.analyze(), meter.analyze(0), len(gdg.resample_time(np.sin(np.arange(5000) * 0.01), 44100, 96000)))
print("sanitize workload done")
