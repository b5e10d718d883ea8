#!/usr/bin/env python3
"""Golden vectors produced by EXECUTING the reference's own Go sources, translated mechanically to Python (go2py.py).

    python tests/golden/go2py/make_reference_vectors.py [/root/reference]

writes tests/golden/translated_reference_vectors.npz (+ .json manifest).  Needs the reference tree, so it runs in the build
container only; the vectors are committed and tests/test_reference_translated.py compares the C oracle (CPU) and the CUDA
product (-m gpu) against them.  Before anything is generated the translated packages must reproduce the reference's own
unit-test vectors (selfcheck.py).

What is executed: effects.CreateUnit / every unit's Process, signal.Chain (AppendUnit / Set* / SetBypass / MoveUp / MoveDown /
RemoveUnit / Process), effects.PreparePowerAmp + the power amp's recompilation, filter.Import / Reduce / Normalize / Multiply /
Add / Process, oversampling, resample, fft, spatializer.Process, level.Meter - all from the .go files under the reference
root, none of it restated by hand.  Stand-ins (shims.py): os.ReadFile, json.Unmarshal, the RIFF container reader."""
import json
import os
import sys
import tempfile
import time
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import load        # noqa: E402
import selfcheck   # noqa: E402
import shims       # noqa: E402
from goruntime import G, GoPanic  # noqa: E402

OUT = os.path.join(HERE, "..", "translated_reference_vectors")
IR_NAMES = ["Guitar: American Vintage (Center)", "Guitar: Custom BP", "PA: HC 4-way (Sub)"]
CHAIN12 = [("noise_gate", {}, {}), ("compressor", {}, {}), ("excess", {"gain": 6}, {"oversampling": "4"}),
           ("fuzz", {}, {"oversampling": "4"}), ("overdrive", {"gain": 20, "drive": 100}, {"oversampling": "4", "valve": "ECC83 (12AX7)"}),
           ("distortion", {"gain": 10, "level": -6}, {"oversampling": "4"}), ("tone_stack", {}, {}),
           ("chorus", {"depth": 100, "speed": 30}, {}), ("delay", {"delay_time": 200, "feedback": -5, "level": -5}, {}),
           ("reverb", {"mix": 50}, {}), ("power_amp", {"level_1": 0}, {"filter_order": "8192", "filter_1": None}), ("cabinet", {}, {})]


def test_signal(fs, n_total, seed, amp=0.7):
    """two partials + noise with a loud / quiet envelope: crosses gate / compressor thresholds and the clamps"""
    rng = np.random.default_rng(seed)
    t = np.arange(n_total) / fs
    x = 0.5 * np.sin(2 * np.pi * 196.0 * t) + 0.25 * np.sin(2 * np.pi * 1234.5 * t + 0.3) + 0.05 * rng.uniform(-1, 1, n_total)
    env = 0.05 + 0.95 * (np.sin(2 * np.pi * 3.0 * t) > -0.3)
    return amp * x * env


def tuner_signal(f0, fs, total):
    """three partials; the test regenerates it from (f0, fs, total) instead of storing 100k samples per case"""
    t = np.arange(total) / fs
    return 0.5 * np.sin(2 * np.pi * f0 * t) + 0.2 * np.sin(2 * np.pi * 2 * f0 * t + 0.4) + 0.1 * np.sin(2 * np.pi * 3 * f0 * t + 1.1)


class Ref(object):
    """the translated reference"""

    def __init__(self, root):
        js = shims.JsonShim()
        self.loader = load.Loader(root, overrides={"os": shims.OsShim(root), "encoding/json": js,
                                                   load.MODULE_PREFIX + "wave": shims.WaveShim()})
        n = selfcheck.run(self.loader, verbose=True)
        self.selfcheck_vectors = n
        self.filter = self.loader.load("filter")
        js.elem_class = self.filter.filterDescriptorStruct
        self.effects = self.loader.load("effects")
        self.signal = self.loader.load("signal")
        self.spatializer = self.loader.load("spatializer")
        self.level = self.loader.load("level")
        self.unit_names = list(self.effects.UnitTypes())
        full = json.load(open(os.path.join(root, "ir", "index.json")))
        idx = [d for d in full if d["Name"] in IR_NAMES]
        assert len(idx) == len(IR_NAMES)
        tf = tempfile.NamedTemporaryFile("w", suffix=".json", delete=False)
        json.dump(idx, tf); tf.close()
        t0 = time.time()
        self.irs, err = self.filter.Import(tf.name)
        os.unlink(tf.name)
        assert err is None, err
        print("filter.Import of %d responses x 7 rates: %.1f s" % (len(idx), time.time() - t0))
        # effects.PreparePowerAmp is what the controller calls on every power amp after creation (controller.go: chains are built
        # with signal.CreateChain(irs), which does it inside AppendUnit)

    def chain(self, spec, ir_name=None):
        ch = self.signal.CreateChain(self.irs)
        for unit, numeric, discrete in spec:
            uid, err = ch.AppendUnit(self.unit_names.index(unit))
            assert err is None, err
            for k, v in discrete.items():
                v = ir_name if v is None else v
                if v is not None:
                    e = ch.SetDiscreteValue(uid, k, v)
                    assert e is None, e
            for k, v in numeric.items():
                e = ch.SetNumericValue(uid, k, int(v))
                assert e is None, e
            assert ch.SetBypass(uid, False) is None
        return ch

    def process(self, ch, x, fs):
        out = G.make_slice(0.0, len(x), None, 'f')
        ch.Process(G.lit([float(v) for v in x], 'f'), out, fs)
        return np.array(out.tolist(), dtype=np.float64)

    def params(self, unit):
        u = self.effects.CreateUnit(self.unit_names.index(unit))
        num, disc = {}, {}
        for p in u.Parameters():
            if p.Type == self.effects.PARAMETER_TYPE_NUMERIC:
                num[p.Name] = (p.Minimum, p.Maximum, p.NumericValue)
            elif p.Type == self.effects.PARAMETER_TYPE_DISCRETE:
                disc[p.Name] = (list(p.DiscreteValues), p.DiscreteValueIndex)
        return num, disc


def unit_variants(ref, name):
    """[(numeric, discrete)]: defaults, all-min, all-max, each discrete value, 2 seeded random draws - the ranges come from
    the translated Parameters() of the unit itself"""
    num, disc = ref.params(name)
    out = [({}, {})]
    if num:
        out.append(({k: v[0] for k, v in num.items()}, {}))
        out.append(({k: v[1] for k, v in num.items()}, {}))
    for k, (vals, _d) in disc.items():
        for v in vals:
            out.append(({}, {k: v}))
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    for _ in range(2):
        out.append(({k: int(rng.integers(v[0], v[1] + 1)) for k, v in num.items()},
                    {k: vals[int(rng.integers(0, len(vals)))] for k, (vals, _d) in disc.items()}))
    return out


def main():
    root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    ref = Ref(root)
    arrays, cases = {}, []
    signals = {}

    def signal(fs, n_total, seed):
        key = "x_%d_%d_%d" % (fs, n_total, seed)
        if key not in signals:
            signals[key] = test_signal(fs, n_total, seed)
            arrays[key] = signals[key]
        return key, signals[key]

    seen = set()

    def stream_case(kind, label, spec, fs, n, blocks, seed, ir_name=None, edits=None, ns=None, amp=1.0, fss=None):
        """ns: block lengths when they vary (then n is the longest and blocks = len(ns)); fss: the sample rate of every block when
        it changes between calls"""
        ident = json.dumps([spec, fs, n, blocks, seed, ir_name, sorted((edits or {}).items()), ns, amp, fss], sort_keys=True, default=list)
        if ident in seen:                                             # e.g. a unit whose only discrete parameter has one value
            return
        seen.add(ident)
        lens = list(ns) if ns else [n] * blocks
        key, x = signal(fs, sum(lens), seed)
        t0 = time.time()
        ch = ref.chain(spec, ir_name)
        ys, pos = [], 0
        for k in range(len(lens)):
            for op in (edits or {}).get(k, ()):                       # chain edits BEFORE block k
                r = getattr(ch, op[0])(*op[1:])
                if isinstance(r, tuple):
                    r = r[-1]
                assert r is None, (op, r)
            ys.append(ref.process(ch, amp * x[pos:pos + lens[k]], fss[k] if fss else fs))
            pos += lens[k]
        y = np.concatenate(ys)
        assert np.all(np.isfinite(y)), label
        cid = "c%04d" % len(cases)
        arrays[cid] = y
        cases.append({"id": cid, "kind": kind, "label": label, "spec": [[u, nu, di] for u, nu, di in spec], "ir": ir_name, "fs": fs, "n": n,
                      "blocks": len(lens), "ns": lens, "fss": list(fss) if fss else [fs] * len(lens), "amp": amp, "input": key, "edits": {str(k): [list(op) for op in v] for k, v in (edits or {}).items()},
                      "peak": float(np.max(np.abs(y)))})
        print("%-6s %-64s %6.1f s  peak %.3g" % (cid, label[:64], time.time() - t0, np.max(np.abs(y))), flush=True)

    # ---- every unit on its own: parameter variants x two rate / block-size combinations --------------------------------
    os_units = ("excess", "fuzz", "overdrive", "distortion")
    for name in ref.unit_names:
        if name == "power_amp":
            continue
        for vi, (num, disc) in enumerate(unit_variants(ref, name)):
            combos = [(48000, 256, 4), (192000, 512, 3)]
            if name in os_units and disc.get("oversampling") in ("2", "4"):
                combos = [(48000, 256, 4)] if vi % 2 else [(96000, 128, 6)]   # the translated FFT is slow: one combination each
            for fs, n, blocks in combos:
                stream_case("unit", "%s %s %s @%d n=%d" % (name, json.dumps(num), json.dumps(disc), fs, n), [(name, num, disc)], fs, n, blocks,
                            seed=zlib.crc32(("%s%d" % (name, fs)).encode()) % 1000)
    # ---- every unit at the other five supported rates (filter.go:25-33), defaults, short blocks ------------------------------------
    for name in ref.unit_names:
        if name == "power_amp":
            continue
        for fs in (22050, 32000, 44100, 88200, 96000):
            if name in ("flanger", "phaser") and fs in (22050, 44100, 88200):
                continue                                              # the reference panics there at depth 100 (recorded below)
            disc = {"oversampling": "2" if fs < 50000 else "4"} if name in os_units else {}
            stream_case("rates", "%s defaults @%d n=128" % (name, fs), [(name, {}, disc)], fs, 128, 5, seed=200 + fs % 97)
    # ---- long-memory units need more blocks than their delay: low rate, many blocks --------------------------------------
    stream_case("unit", "reverb mix 70, taps reached", [("reverb", {"mix": 70}, {})], 22050, 512, 20, 5)
    stream_case("unit", "delay 30 ms feedback", [("delay", {"delay_time": 30, "feedback": -3, "level": 0}, {})], 32000, 256, 12, 6)
    stream_case("unit", "noise gate hold", [("noise_gate", {"threshold_open": -18, "threshold_close": -30, "hold_time": 20}, {})], 22050, 512, 20, 7)
    for spec in ([("chorus", {"depth": 100, "speed": 100}, {})], [("auto_yoy", {"depth": 100}, {"follow": "envelope"})],
                 [("flanger", {"depth": 100, "speed": 100}, {})], [("phaser", {"depth": 100, "speed": 100, "phase": -90}, {})]):
        stream_case("unit", "%s short blocks" % spec[0][0], spec, 48000, 32, 40, 9)
    # ---- power amp: mixes, Reduce, recompilation --------------------------------------------------------------------------
    pa = "Guitar: American Vintage (Center)"
    stream_case("power_amp", "one response, default order", [("power_amp", {"level_1": -3}, {"filter_1": pa})], 96000, 1024, 4, 11)
    stream_case("power_amp", "two responses mixed", [("power_amp", {"level_1": -3, "level_2": -9}, {"filter_1": pa, "filter_2": "Guitar: Custom BP"})], 96000, 1024, 3, 12)
    stream_case("power_amp", "filter.Reduce to 2048", [("power_amp", {"level_1": 0}, {"filter_order": "2048", "filter_1": "Guitar: Custom BP"})], 96000, 1024, 3, 13)
    stream_case("power_amp", "filter.Reduce to 512 at 48 kHz", [("power_amp", {"level_1": 0}, {"filter_order": "512", "filter_1": pa})], 48000, 512, 4, 14)
    stream_case("power_amp", "no response: silence", [("power_amp", {}, {})], 48000, 256, 2, 15)
    stream_case("power_amp", "recompiled between blocks", [("power_amp", {"level_1": 0}, {"filter_order": "1024", "filter_1": pa})], 96000, 512, 6, 16,
                edits={2: [("SetNumericValue", 0, "level_1", -12)], 4: [("SetDiscreteValue", 0, "filter_2", "Guitar: Custom BP")]})
    stream_case("power_amp", "long response partitioned at a JACK period (HC 4-way Sub, 9599 taps at 96 kHz)",
                [("power_amp", {"level_1": -6}, {"filter_1": "PA: HC 4-way (Sub)"})], 96000, 1024, 12, 17)
    for fs in (22050, 32000, 44100, 48000, 88200, 96000, 192000):     # filter.Import's resampling at every supported rate
        stream_case("power_amp", "Custom BP reduced to 256 @%d" % fs, [("power_amp", {"level_1": -2}, {"filter_order": "256", "filter_1": "Guitar: Custom BP"})],
                    fs, 256, 3, 18)
    # ---- chains -----------------------------------------------------------------------------------------------------------
    stream_case("chain", "CHAIN12 @96 kHz n=1024", CHAIN12, 96000, 1024, 4, 21, ir_name=pa)
    stream_case("chain", "CHAIN12 @192 kHz n=2048", CHAIN12, 192000, 2048, 2, 22, ir_name=pa)
    every = [(u, {}, {}) for u in ref.unit_names if u != "power_amp"]
    stream_case("chain", "all 20 units in menu order, defaults", every, 48000, 256, 4, 23)
    stream_case("chain", "bypassed units and an empty chain are pass-through", [], 48000, 128, 2, 24)
    base = [("distortion", {"gain": 8}, {}), ("tremolo", {}, {}), ("delay", {"delay_time": 10}, {}), ("tone_stack", {}, {})]
    stream_case("chain", "edits between blocks: move, bypass, remove, append", base, 48000, 256, 8, 25,
                edits={1: [("MoveDown", 0)], 2: [("SetBypass", 2, True)], 3: [("MoveUp", 3)], 4: [("RemoveUnit", 1)],
                       5: [("AppendUnit", ref.unit_names.index("ring_modulator")), ("SetBypass", 3, False)], 6: [("SetBypass", 1, False)]})

    # ---- block length changes between calls (the up-sampler's history is re-made, the decimator's tail survives: oversampling.go:74-84)
    for name in os_units:
        for f in ("2", "4"):
            stream_case("varn", "%s x%s, block length changes" % (name, f), [(name, {"gain": 9}, {"oversampling": f})], 96000, 512, 0, 61,
                        ns=[256, 256, 512, 128, 128, 256, 64, 512])
    stream_case("varn", "delay / chorus / reverb / tone_stack, block length changes",
                [("delay", {"delay_time": 4}, {}), ("chorus", {}, {}), ("reverb", {}, {}), ("tone_stack", {}, {})], 48000, 1024, 0, 62,
                ns=[128, 1024, 256, 256, 64, 512, 32, 1024])
    stream_case("varn", "power amp (short response), block length changes", [("power_amp", {"level_1": -3}, {"filter_order": "512", "filter_1": pa})],
                48000, 512, 0, 63, ns=[256, 512, 128, 128, 512, 64])

    # ---- the sample rate changes between calls (units re-derive coefficients, delay-type units re-make their histories) ----------
    rates_seq = [48000] * 3 + [96000] * 3 + [48000] * 2 + [192000] * 2
    for spec in ([("delay", {"delay_time": 3}, {}), ("chorus", {}, {})], [("reverb", {"mix": 60}, {})], [("flanger", {}, {}), ("phaser", {}, {})],
                 [("tone_stack", {}, {}), ("cabinet", {}, {}), ("bandpass", {}, {}), ("auto_wah", {}, {})],
                 [("overdrive", {"gain": 10}, {"oversampling": "2"}), ("fuzz", {}, {"oversampling": "4"})],
                 [("compressor", {}, {}), ("tremolo", {}, {}), ("ring_modulator", {}, {}), ("signal_generator", {"signal_amplitude": 20}, {})],
                 [("noise_gate", {}, {}), ("octaver", {}, {})], [("auto_yoy", {}, {})],
                 [("power_amp", {"level_1": -3}, {"filter_order": "1024", "filter_1": pa})]):
        stream_case("ratechange", "%s, sample rate changes" % "+".join(u[0] for u in spec), spec, 48000, 256, 0, 71, ns=[256] * len(rates_seq), fss=rates_seq)

    # ---- random chains: units, parameter values, block length and rate drawn at random (seeded) ------------------------------
    rates = [22050, 32000, 44100, 48000, 88200, 96000, 192000]
    schema = {u: ref.params(u) for u in ref.unit_names}
    for seed in range(48):
        rng = np.random.default_rng(7000 + seed)
        fs = int(rng.choice(rates))
        n = int(rng.choice([64, 256, 1024, 2048]))
        spec = []
        for _ in range(int(rng.integers(2, 9))):
            u = str(rng.choice(ref.unit_names))
            num, disc = schema[u]
            numeric, discrete = {}, {}
            for k, (mn, mx, _d) in num.items():
                if rng.random() >= 0.35:
                    numeric[k] = int(rng.integers(mn, mx + 1))
            for k, (vals, _d) in disc.items():
                if rng.random() < 0.35:
                    continue
                if u == "power_amp" and k == "filter_order":
                    vals = [v for v in vals if int(v) <= 8192]
                if u == "power_amp" and k.startswith("filter_") and k != "filter_order":
                    vals = ["- NONE -"] + IR_NAMES[:2]
                discrete[k] = str(rng.choice(vals))
            if u == "power_amp" and "filter_order" not in discrete:
                discrete["filter_order"] = "8192"
            if u == "power_amp" and all(discrete.get("filter_%d" % i, "- NONE -") == "- NONE -" for i in range(1, 9)):
                discrete["filter_1"] = IR_NAMES[0]                     # a power amp without a response is silence (covered above)
            spec.append((u, numeric, discrete))
        if fs in (22050, 44100, 88200) and any(u[0] in ("flanger", "phaser") for u in spec):
            fs = 48000                                                 # the reference panics there at depth 100 (see below)
        # discontinuous units where their input is bit-identical (tests/test_random_chains.py): gates, then one octaver / auto_yoy
        gates = [u for u in spec if u[0] == "noise_gate"]
        octs = [u for u in spec if u[0] in ("octaver", "auto_yoy")][:1]
        spec = gates + octs + [u for u in spec if u[0] not in ("noise_gate", "octaver", "auto_yoy")]
        blocks = 3 if n >= 1024 else 5
        stream_case("random", "seed %d: %s @%d n=%d" % (seed, "+".join(u[0] for u in spec), fs, n), spec, fs, n, blocks, 100 + seed,
                    amp=float(rng.choice([0.15, 1.0, 1.4])))

    # ---- spatializer ------------------------------------------------------------------------------------------------------
    C, fs, n, blocks = 5, 96000, 512, 4
    sp = ref.spatializer.Create(C)
    sp.SetSampleRate(fs)
    geo = []
    for c in range(C):
        az, dist, lvl = -80.0 + 40.0 * c, 0.7 + 0.45 * c, 1.0 - 0.15 * c
        assert sp.SetAzimuth(c, az) is None and sp.SetDistance(c, dist) is None and sp.SetLevel(c, lvl) is None
        geo.append([az, dist, lvl])
    xs = np.stack([test_signal(fs, n * blocks, 30 + c, 0.5) for c in range(C)])
    aux = 0.2 * test_signal(fs, n * blocks, 40, 0.5)
    outs = []
    for k in range(blocks):
        ins = G.lit([G.lit([float(v) for v in xs[c, k * n:(k + 1) * n]], 'f') for c in range(C)])
        a = G.lit([float(v) for v in aux[k * n:(k + 1) * n]], 'f') if k % 2 == 0 else None
        o = G.lit([G.make_slice(0.0, n, None, 'f'), G.make_slice(0.0, n, None, 'f')])
        sp.Process(ins, a, o)
        outs.append(np.array([o[0].tolist(), o[1].tolist()]))
    arrays["spat_in"], arrays["spat_aux"], arrays["spat_out"] = xs, aux, np.concatenate(outs, axis=1)
    spat = {"channels": C, "fs": fs, "n": n, "blocks": blocks, "geometry": geo, "aux_on_even_blocks": True, "other_rates": [192000, 44100]}
    for rate in spat["other_rates"]:                                   # SetSampleRate never stores the rate (spatializer.go:418-431): observed here
        sp2 = ref.spatializer.Create(C)
        sp2.SetSampleRate(rate)
        for c, (az, dist, lvl) in enumerate(geo):
            assert sp2.SetAzimuth(c, az) is None and sp2.SetDistance(c, dist) is None and sp2.SetLevel(c, lvl) is None
        outs2 = []
        for k in range(blocks):
            ins = G.lit([G.lit([float(v) for v in xs[c, k * n:(k + 1) * n]], 'f') for c in range(C)])
            o = G.lit([G.make_slice(0.0, n, None, 'f'), G.make_slice(0.0, n, None, 'f')])
            sp2.Process(ins, None, o)
            outs2.append(np.array([o[0].tolist(), o[1].tolist()]))
        arrays["spat_out_%d" % rate] = np.concatenate(outs2, axis=1)
    print("spatializer: %d channels x %d blocks" % (C, blocks))

    # ---- IR algebra: what filter.Import resampled, and Reduce / Normalize / Multiply / Add on it ---------------------------
    ir_info = []
    for name in IR_NAMES[:2]:
        for rate in (44100, 96000, 192000):
            f = ref.filter if False else ref.irs.CreateFilter(name, rate)
            key = "ir_%d_%d" % (IR_NAMES.index(name), rate)
            arrays[key] = np.array(f.Coefficients().tolist())
            ir_info.append({"name": name, "rate": rate, "key": key})
    f1 = ref.irs.CreateFilter(IR_NAMES[0], 96000)
    f2 = ref.irs.CreateFilter(IR_NAMES[1], 96000)
    arrays["alg_reduce_1024"] = np.array(f2.Reduce(1024).Coefficients().tolist())
    arrays["alg_normalize"] = np.array(f1.Normalize().Coefficients().tolist())
    arrays["alg_multiply"] = np.array(f1.Multiply(0.5).Coefficients().tolist())
    s, err = f1.Normalize().Add(f2.Normalize())
    assert err is None
    arrays["alg_add"] = np.array(s.Coefficients().tolist())

    # ---- level meter (level.Meter.Process / Analyze), streaming -----------------------------------------------------------
    fs, n, blocks = 8000, 2048, 12                                   # 2 s hold = 16000 samples: the peak hold expires inside the run
    m, err = ref.level.CreateMeter(2, G.lit(["a", "b"]))
    assert err is None
    m.SetEnabled(True)
    xm = np.stack([test_signal(fs, n * blocks, 50, 0.9) * np.linspace(1.0, 0.02, n * blocks), 0.3 * test_signal(fs, n * blocks, 51, 0.9)])
    reads = []
    for k in range(blocks):
        m.Process(G.lit([G.lit([float(v) for v in xm[c, k * n:(k + 1) * n]], 'f') for c in range(2)]), fs)
        row = []
        for c in range(2):
            r, err = m.Analyze(c)
            assert err is None
            row.append([int(r.Level()), int(r.Peak())])
        reads.append(row)
    arrays["meter_in"] = xm
    meter = {"fs": fs, "n": n, "blocks": blocks, "readouts": reads}
    print("level meter read-outs:", reads[-1])

    # ---- parameter tables (effects.Unit.Parameters through signal.Chain, the power amp after PreparePowerAmp) and error strings ----
    ch = ref.signal.CreateChain(ref.irs)
    tables = []
    for t in range(len(ref.unit_names)):
        uid, err = ch.AppendUnit(t)
        assert err is None
        ps, err = ch.Parameters(uid)
        assert err is None
        tables.append([{"Name": p.Name, "Type": int(p.Type), "PhysicalUnit": p.PhysicalUnit, "Minimum": int(p.Minimum), "Maximum": int(p.Maximum),
                        "NumericValue": int(p.NumericValue), "DiscreteValueIndex": int(p.DiscreteValueIndex),
                        "DiscreteValues": list(p.DiscreteValues) if p.DiscreteValues is not None else []} for p in ps])
        assert ch.UnitType(uid)[0] == t and ch.GetBypass(uid)[0] is True          # appended bypassed (signal.go:74)
    ch = ref.signal.CreateChain(ref.irs)
    od, _ = ch.AppendUnit(ref.unit_names.index("overdrive"))

    def msg(r):
        e = r[-1] if isinstance(r, tuple) else r
        assert e is not None
        return e.Error()
    errors = {
        "set_numeric_value out of range": msg(ch.SetNumericValue(od, "gain", 31)),
        "set_numeric_value unknown": msg(ch.SetNumericValue(od, "nope", 1)),
        "set_numeric_value on discrete": msg(ch.SetNumericValue(od, "valve", 1)),
        "set_discrete_value bad value": msg(ch.SetDiscreteValue(od, "valve", "EL34")),
        "set_discrete_value on numeric": msg(ch.SetDiscreteValue(od, "gain", "x")),
        "get_discrete_value on numeric": msg(ch.GetDiscreteValue(od, "gain")),
        "get_numeric_value unknown": msg(ch.GetNumericValue(od, "nope")),
        "set_numeric_value no unit": msg(ch.SetNumericValue(7, "gain", 1)),
        "set_bypass true no unit": msg(ch.SetBypass(3, True)),
        "set_bypass false no unit": msg(ch.SetBypass(3, False)),
        "get_bypass no unit": msg(ch.GetBypass(3)),
        "remove_unit -1": msg(ch.RemoveUnit(-1)),
        "move_up 0": msg(ch.MoveUp(0)),
        "move_down 0": msg(ch.MoveDown(0)),
        "unit_type no unit": msg(ch.UnitType(5)),
        "append_unit 21": msg(ch.AppendUnit(21)),
        "parameters no unit": msg(ch.Parameters(9)),
    }
    assert ch.GetNumericValue(od, "gain")[0] == 0                              # a rejected value leaves the parameter unchanged
    print("parameter tables of %d units, %d error strings" % (len(tables), len(errors)))

    # ---- tuner.Process / Analyze on synthetic tones (the six recordings of tuner_test.go are in the self-check / oracle goldens) ----
    tn = ref.loader.load("tuner")
    tuner_cases = []
    for f0, fs, n in ((82.4069, 96000, 8192), (110.0, 48000, 4096), (329.63, 44100, 2048), (246.94, 192000, 8192)):
        nb = (96000 + 3 * n) // n                                                # more than one ring of 96000 samples
        x = tuner_signal(f0, fs, nb * n)
        tu = tn.Create()
        for k in range(nb):
            tu.Process(G.lit([float(v) for v in x[k * n:(k + 1) * n]], 'f'), fs)
        t0 = time.time()
        res, err = tu.Analyze()
        assert err is None, err
        tuner_cases.append({"blocks": nb, "fs": fs, "n": n, "f0": f0, "note": res.Note(), "cents": int(res.Cents()), "frequency": float(res.Frequency())})
        print("tuner %.2f Hz @%d: %s %+d cents %.4f Hz (%.1f s)" % (f0, fs, res.Note(), res.Cents(), res.Frequency(), time.time() - t0))

    # ---- BASELINE.json config 1 in miniature: PCM16 payload -> overdrive -> tone_stack -> PCM16 payload, 48 kHz, blocks of 8192 with
    # the last one zero-padded (controller.go:3022-3107), every step the reference's own code: wave.go's conversions and the chain
    wv = ref.loader.load("wave")
    fs1, total = 48000, 24000
    x1 = test_signal(fs1, total, 77, 0.9)
    pcm_in, err = wv.samplesToBytesLPCM16(G.lit([float(v) for v in x1], 'f'))
    assert err is None
    smp, err = wv.bytesToSamplesLPCM16(pcm_in)
    assert err is None
    ch1 = ref.chain([("overdrive", {"gain": 12}, {}), ("tone_stack", {}, {})])
    nblk = (total + 8191) // 8192
    padded = np.zeros(nblk * 8192); padded[:total] = np.array(smp.tolist())
    y1 = np.concatenate([ref.process(ch1, padded[k * 8192:(k + 1) * 8192], fs1) for k in range(nblk)])
    pcm_out, err = wv.samplesToBytesLPCM16(G.lit([float(v) for v in y1], 'f'))
    assert err is None
    arrays["cfg1_pcm_in"] = np.array(pcm_in.tolist(), dtype=np.uint8)
    arrays["cfg1_pcm_out"] = np.array(pcm_out.tolist(), dtype=np.uint8)
    cfg1 = {"fs": fs1, "samples": total, "block": 8192, "blocks": nblk, "spec": [["overdrive", {"gain": 12}, {}], ["tone_stack", {}, {}]]}
    print("config 1: %d PCM16 samples -> %d blocks -> %d bytes" % (total, nblk, len(pcm_out)))

    # ---- the reference's out-of-range read: flanger / phaser at 22.05 / 44.1 / 88.2 kHz with depth 100 --------
    panics = []
    for unit, fs, n in (("flanger", 44100, 64), ("phaser", 22050, 32), ("flanger", 88200, 128), ("flanger", 48000, 64)):
        ch = ref.chain([(unit, {"depth": 100, "speed": 37}, {})])
        x = test_signal(fs, n * 600, 3)
        hit = None
        for k in range(600):
            try:
                ref.process(ch, x[k * n:(k + 1) * n], fs)
            except GoPanic as e:
                hit = {"block": k, "message": str(e)}
                break
        panics.append({"unit": unit, "fs": fs, "n": n, "panic": hit})
        print("%s @%d n=%d: %s" % (unit, fs, n, hit))

    np.savez_compressed(OUT + ".npz", **arrays)
    manifest = {"generator": "tests/golden/go2py/make_reference_vectors.py", "reference": "andrepxx/go-dsp-guitar v1.8.0 (the tree under /root/reference)",
                "selfcheck_vectors": ref.selfcheck_vectors, "ir_names": IR_NAMES, "unit_names": ref.unit_names, "cases": cases, "spatializer": spat, "ir": ir_info,
                "level_meter": meter, "reference_panics": panics, "config1": cfg1, "parameters": tables, "errors": errors, "tuner": tuner_cases, "integer_division_sites": sorted(G.intdiv_sites)}
    json.dump(manifest, open(OUT + ".json", "w"), indent=0)
    print("%d stream cases, %d arrays, %.1f MB" % (len(cases), len(arrays), os.path.getsize(OUT + ".npz") / 1e6))
    print("integer-division sites executed:", sorted(G.intdiv_sites))


if __name__ == "__main__":
    main()
