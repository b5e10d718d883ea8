#!/usr/bin/env python3
"""Random chains (1 - 6 units, parameters at random incl. their extremes, any supported rate, fixed or varying block lengths)
through the EXECUTED reference (go2py) and the C oracle, compared block by block at 1e-12 of the peak.  Needs the reference
tree.    python tests/golden/go2py/hunt.py FIRST_SEED LAST_SEED
4400 seeds: 0 mismatches, no reference panic (round 2)."""
import sys, os, json, time
HERE = os.path.dirname(os.path.abspath(__file__))
for p_ in (HERE, os.path.join(HERE, '..', '..'), os.path.join(HERE, '..', '..', '..', 'oracle')):
    sys.path.insert(0, p_)
import numpy as np
import common
from common import orc
import make_reference_vectors as gen
from goruntime import GoPanic
def hunt(lo, hi, root='/root/reference', ref=None, verbose=True):
    """returns the list of (seed, worst error, label) that differ; reference panics are reported and skipped"""
    ref = ref or gen.Ref(root)
    z = np.load(os.path.join(HERE, '..', 'ir_bank.npz'))
    ob = orc.IRBank()
    for k, name in enumerate(z["names"]):
        if str(name) in gen.IR_NAMES:
            ob.add(str(name), int(z["compensation"][k]), orc.pcm16_decode(z["pcm_%02d" % k]), int(z["sample_rate"]))
    rates = [22050, 32000, 44100, 48000, 88200, 96000, 192000]
    schema = {u: ref.params(u) for u in ref.unit_names}
    bad = []; t0=time.time()
    for seed in range(lo, hi):
        rng = np.random.default_rng(90000 + seed)
        fs = int(rng.choice(rates)); n = int(rng.choice([32, 64, 256, 512, 1024]))
        spec = []
        for _ in range(int(rng.integers(1, 7))):
            u = str(rng.choice(ref.unit_names))
            num, disc = schema[u]
            numeric, discrete = {}, {}
            for k, (mn, mx, _d) in num.items():
                r = rng.random()
                if r < 0.1: numeric[k] = mn
                elif r < 0.2: numeric[k] = mx
                elif r < 0.7: numeric[k] = int(rng.integers(mn, mx + 1))
            for k, (vals, _d) in disc.items():
                if rng.random() < 0.4: continue
                if u == "power_amp" and k == "filter_order": vals = [v for v in vals if int(v) <= 4096]
                if u == "power_amp" and k.startswith("filter_") and k != "filter_order": vals = ["- NONE -"] + gen.IR_NAMES[:2]
                discrete[k] = str(rng.choice(vals))
            if u == "power_amp":
                discrete.setdefault("filter_order", "1024")
                if all(discrete.get("filter_%d" % i, "- NONE -") == "- NONE -" for i in range(1, 9)): discrete["filter_1"] = gen.IR_NAMES[0]
            spec.append((u, numeric, discrete))
        if fs in (22050, 44100, 88200) and any(u[0] in ("flanger", "phaser") for u in spec): fs = 48000
        blocks = int(rng.integers(3, 7))
        # block lengths may vary, too
        ns = [n] * blocks if rng.random() < 0.7 else [int(rng.choice([32, 64, 128, 256, 512])) for _ in range(blocks)]
        x = gen.test_signal(fs, sum(ns), seed) * float(rng.choice([0.1, 1.0, 1.6]))
        ch = ref.chain(spec); oc = common.build_chain(orc.Chain(ob), spec)
        pos = 0; worst = 0.0; peak = 0.0
        try:
            for nb in ns:
                yr = ref.process(ch, x[pos:pos+nb], fs); yo = oc.process(np.ascontiguousarray(x[pos:pos+nb]), fs); pos += nb
                worst = max(worst, float(np.max(np.abs(yr-yo)))); peak = max(peak, float(np.max(np.abs(yr))))
        except GoPanic as e:
            print("seed", seed, "reference panic:", e, [s[0] for s in spec], fs); continue
        if not (worst <= 1e-12 * max(1.0, peak)):
            bad.append((seed, worst, json.dumps(spec)))
            print("seed %d MISMATCH %.3e peak %.3g fs %d ns %s spec %s" % (seed, worst, peak, fs, ns, json.dumps(spec)))
    if verbose:
        print("seeds %d..%d: %d mismatches, %.0f s" % (lo, hi, len(bad), time.time()-t0))
    return bad


if __name__ == "__main__":
    sys.exit(1 if hunt(int(sys.argv[1]), int(sys.argv[2])) else 0)
