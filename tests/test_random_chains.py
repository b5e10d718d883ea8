"""Random chains: unit sequences, parameter values, block lengths and sample rates drawn at random (seeded), each chain
streamed for a few blocks against the oracle.  Finds interactions the hand-written cases do not think of."""
import numpy as np
import pytest

import common
from common import gdg, orc

BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]
RATES = [22050, 32000, 44100, 48000, 88200, 96000, 192000]


@pytest.fixture(params=BACKENDS)
def lib(request):
    return common.backend_lib(request.param)


def _schema(lib):
    """{unit: [parameter dicts]} through the product's own Parameters() (checked against the reference tables elsewhere)"""
    out = {}
    probe = gdg.Chain(lib=lib)
    for u in gdg.UNIT_TYPES:
        i = probe.append_unit(u)
        out[u] = probe.parameters(i)
    return out


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("GDG_RANDOM_CHAINS", "48"))))
def test_random_chain_streams_like_the_oracle(lib, seed):
    rng = np.random.default_rng(1000 + seed)
    schema = _schema(lib)
    banks = common.make_banks(lib)
    ir_names = list(np.load(common.GOLD + "/ir_bank.npz")["names"])
    fs = int(rng.choice(RATES))
    n = int(rng.choice([64, 256, 1024, 2048, 8192]))
    spec = []
    for _ in range(int(rng.integers(2, 9))):
        u = str(rng.choice(gdg.UNIT_TYPES))
        numeric, discrete = {}, {}
        for p in schema[u]:
            if rng.random() < 0.35:
                continue                                         # leave at its default
            if p["DiscreteValues"]:
                vals = p["DiscreteValues"]
                if u == "power_amp" and p["Name"] == "filter_order":
                    vals = [v for v in vals if int(v) <= (65536 if n == 8192 else 8192)]   # longer filters need 8192-sample blocks
                if u == "power_amp" and p["Name"].startswith("filter_") and p["Name"] != "filter_order":
                    vals = ["- NONE -"] + [str(x) for x in rng.choice(ir_names, 2)]
                discrete[p["Name"]] = str(rng.choice(vals))
            else:
                numeric[p["Name"]] = int(rng.integers(int(p["Minimum"]), int(p["Maximum"]) + 1))
        if u == "power_amp" and "filter_order" not in discrete:
            discrete["filter_order"] = "8192" if n != 8192 else "65536"                 # the default (1048576) needs n = 8192
        spec.append((u, numeric, discrete))
    # discontinuous units (gate open/close, octaver polarity flips with hysteresis, auto_yoy's clamp edges) amplify a 1e-18 difference of their
    # input into a different decision - between Go and any restatement too.  Like CHAIN12 the random chains
    # keep them where their input is bit-identical: gates first, then at most one octaver, then everything else.
    # auto_yoy belongs here too: where its delay is clamped to a whole number of samples both interpolation weights are 1
    # (autoyoy.go:124-140), a hair off it they are (1, 0) - the output jumps by a whole sample at the clamp edge.
    gates = [u for u in spec if u[0] == "noise_gate"]
    octs = [u for u in spec if u[0] in ("octaver", "auto_yoy")][:1]
    spec = gates + octs + [u for u in spec if u[0] not in ("noise_gate", "octaver", "auto_yoy")]
    a, b = common.make_chain_pair(lib, spec, banks)
    blocks = 3 if n >= 2048 else 6
    x = common.synth_fast(seed, fs, blocks * n, seed) * float(rng.choice([0.1, 1.0, 1.7]))
    # a random chain may be silent by construction (a power amp without a filter, a gate that never opens): not asserted here
    common.compare_stream(a, b, x, n, fs, label="seed %d fs %d n %d %s" % (seed, fs, n, [s[0] for s in spec]), expect_signal=False)


def _random_params(rng, schema, u, n, ir_names):
    numeric, discrete = {}, {}
    for p in schema[u]:
        if rng.random() < 0.5:
            continue
        if p["DiscreteValues"]:
            vals = p["DiscreteValues"]
            if u == "power_amp" and p["Name"] == "filter_order":
                vals = [v for v in vals if int(v) <= 8192]
            if u == "power_amp" and p["Name"].startswith("filter_") and p["Name"] != "filter_order":
                vals = ["- NONE -"] + [str(x) for x in rng.choice(ir_names, 2)]
            discrete[p["Name"]] = str(rng.choice(vals))
        else:
            numeric[p["Name"]] = int(rng.integers(int(p["Minimum"]), int(p["Maximum"]) + 1))
    if u == "power_amp" and "filter_order" not in discrete:
        discrete["filter_order"] = "4096"
    return numeric, discrete


CONTINUOUS = [u for u in gdg.UNIT_TYPES if u not in ("noise_gate", "octaver", "auto_yoy")]


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("GDG_RANDOM_EDITS", "24"))))
def test_random_edit_sequences(lib, seed):
    """A live chain is edited between blocks the way the web UI drives signal.Chain (append / remove / move / bypass /
    parameter changes, a new block length, a new sample rate): unit state has to travel, reset or be rebuilt exactly
    where the reference does it (signal.go:52-218, each unit's Process head, poweramp.go:132-181)."""
    rng = np.random.default_rng(5000 + seed)
    schema = _schema(lib)
    banks = common.make_banks(lib)
    ir_names = list(np.load(common.GOLD + "/ir_bank.npz")["names"])
    a, b = orc.Chain(banks[0]), gdg.Chain(banks[1], lib=lib)
    fs, n = int(rng.choice(RATES)), int(rng.choice([256, 1024, 2048]))

    def both(fn):
        ra, rb = fn(a), fn(b)
        return ra, rb

    def append():
        u = str(rng.choice(CONTINUOUS))
        ia, ib = both(lambda c: c.append_unit(u))
        assert ia == ib
        numeric, discrete = _random_params(rng, schema, u, n, ir_names)
        for k, v in discrete.items():
            both(lambda c: c.set_discrete_value(ia, k, v))
        for k, v in numeric.items():
            both(lambda c: c.set_numeric_value(ia, k, v))
        both(lambda c: c.set_bypass(ia, False))
        kinds.append(u)

    kinds = []
    for _ in range(int(rng.integers(2, 6))):
        append()
    pos = 0
    x = common.synth_fast(seed, 48000, 40 * 2048, seed)      # just a long varied signal; the rate label is what changes
    for step in range(14):
        op = rng.choice(["none", "param", "bypass", "move_up", "move_down", "remove", "append", "block", "rate"],
                        p=[0.15, 0.25, 0.1, 0.08, 0.08, 0.08, 0.1, 0.08, 0.08])
        L = len(kinds)
        if op == "param" and L:
            i = int(rng.integers(L))
            numeric, discrete = _random_params(rng, schema, kinds[i], n, ir_names)
            for k, v in discrete.items():
                both(lambda c: c.set_discrete_value(i, k, v))
            for k, v in numeric.items():
                both(lambda c: c.set_numeric_value(i, k, v))
        elif op == "bypass" and L:
            i = int(rng.integers(L)); v = bool(rng.integers(2))
            both(lambda c: c.set_bypass(i, v))
        elif op == "move_up" and L > 1:
            i = int(rng.integers(1, L)); both(lambda c: c.move_up(i)); kinds[i - 1], kinds[i] = kinds[i], kinds[i - 1]
        elif op == "move_down" and L > 1:
            i = int(rng.integers(0, L - 1)); both(lambda c: c.move_down(i)); kinds[i + 1], kinds[i] = kinds[i], kinds[i + 1]
        elif op == "remove" and L > 1:
            i = int(rng.integers(L)); both(lambda c: c.remove_unit(i)); kinds.pop(i)
        elif op == "append" and L < 9:
            append()
        elif op == "block":
            n = int(rng.choice([64, 256, 1024, 2048, 4096]))
        elif op == "rate":
            fs = int(rng.choice(RATES))
        blk = x[pos:pos + n]; pos += n
        ya, yb = a.process(blk, fs), b.process(blk, fs)
        assert a.length() == b.length() == len(kinds)
        e = common.rms(ya - yb)
        assert np.all(np.isfinite(yb)) and e <= common.TOL_RMS, "seed %d step %d after %s: rms %.3e, chain %s, n %d fs %d" % (seed, step, op, e, kinds, n, fs)


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("GDG_RANDOM_MIX", "16"))))
def test_random_spatializer_and_power_amp_sessions(lib, seed):
    """Spatializer positions / levels and the power amp's eight filter slots change between blocks, block length and rate
    too: the mix-down keeps its inter-aural history, the power amp recompiles with a fresh overlap tail
    (poweramp.go:132-181) exactly when the reference does."""
    rng = np.random.default_rng(9000 + seed)
    banks = common.make_banks(lib)
    ir_names = [str(x) for x in np.load(common.GOLD + "/ir_bank.npz")["names"]]
    C = int(rng.integers(1, 41))
    sa, sb = orc.Spatializer(C), gdg.Spatializer(C, lib=lib)
    a, b = common.make_chain_pair(lib, [("power_amp", {}, {"filter_order": "2048"})], banks)
    fs, n = int(rng.choice(RATES)), int(rng.choice([256, 1024, 4096]))
    x = np.stack([common.synth_fast(c, 48000, 12 * 4096, seed) for c in range(min(C, 6))])
    pos = 0
    for step in range(10):
        for _ in range(int(rng.integers(0, 4))):
            c, what = int(rng.integers(C)), rng.choice(["az", "dist", "level"])
            v = float(rng.uniform(-90, 90)) if what == "az" else (float(rng.uniform(0.0, 10.0)) if what == "dist" else float(rng.uniform(0.0, 1.0)))
            for s in (sa, sb):
                {"az": s.set_azimuth, "dist": s.set_distance, "level": s.set_level}[what](c, v)
        r = rng.random()
        if r < 0.2:
            fs = int(rng.choice(RATES)); sa.set_sample_rate(fs); sb.set_sample_rate(fs)
        elif r < 0.4:
            n = int(rng.choice([64, 256, 1024, 4096]))
        elif r < 0.8:
            k = int(rng.integers(1, 9))
            what = rng.choice(["filter", "level", "order"])
            for ch in (a, b):
                if what == "filter":
                    ch.set_discrete_value(0, "filter_%d" % k, str(rng.choice(["- NONE -"] + ir_names)) if ch is a else last)
                elif what == "level":
                    ch.set_numeric_value(0, "level_%d" % k, int(rng.integers(-60, 1)) if ch is a else last)
                else:
                    ch.set_discrete_value(0, "filter_order", str(rng.choice(["64", "512", "2048", "8192"])) if ch is a else last)
                last = ch.get_discrete_value(0, "filter_%d" % k) if what == "filter" else (ch.get_numeric_value(0, "level_%d" % k) if what == "level" else ch.get_discrete_value(0, "filter_order"))
        blk = np.stack([x[c % x.shape[0], pos:pos + n] * (1.0 + 0.1 * (c // x.shape[0])) for c in range(C)]); pos += n
        aux = 0.05 * x[0, pos:pos + n] if rng.random() < 0.3 else None
        ya, yb = sa.process(blk, aux), sb.process(blk, aux)
        scale = max(1.0, float(np.max(np.abs(ya))))
        assert np.max(np.abs(ya - yb)) <= 1e-11 * scale, (seed, step, C, n, fs)
        pa, pb = a.process(blk[0], fs), b.process(blk[0], fs)
        assert common.rms(pa - pb) <= common.TOL_RMS, (seed, step, n, fs)


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("GDG_RANDOM_SESSIONS", "10"))))
def test_random_controller_sessions(lib, seed):
    """controller.process as a whole on random batches: 1..9 channels whose chains differ in length (so the batch mixes
    segment rounds), bypass toggles and tuner-channel changes between blocks, tuner read-outs at random moments, the level
    meter on, blocking calls and submit / wait pairs interleaved."""
    rng = np.random.default_rng(13000 + seed)
    schema = _schema(lib)
    banks = common.make_banks(lib)
    ir_names = list(np.load(common.GOLD + "/ir_bank.npz")["names"])
    C = int(rng.integers(1, 10))
    fs, n = int(rng.choice([44100, 48000, 96000, 192000])), int(rng.choice([256, 1024, 2048]))
    oc, gc = [], []
    for c in range(C):
        spec = []
        for _ in range(int(rng.choice([1, 3, 6, 27]))):
            u = str(rng.choice([k for k in CONTINUOUS if k != "power_amp"]))
            numeric, discrete = _random_params(rng, schema, u, n, ir_names)
            spec.append((u, numeric, discrete))
        a, b = common.make_chain_pair(lib, spec, banks)
        oc.append(a); gc.append(b)
    sa, sb = orc.Spatializer(C), gdg.Spatializer(C, lib=lib)
    ta, tb = orc.Tuner(), gdg.Tuner(lib=lib)
    ma, mb = orc.Meter(2 * C + 3), gdg.Meter(2 * C + 3, lib=lib)
    ma.set_enabled(True); mb.set_enabled(True)
    x = np.stack([common.synth_fast(c, fs, 10 * n, seed) for c in range(C)])
    tch = int(rng.integers(-1, C))
    pending = None
    for k in range(10):
        if rng.random() < 0.3:
            c = int(rng.integers(C)); i = int(rng.integers(oc[c].length())); v = bool(rng.integers(2))
            oc[c].set_bypass(i, v); gc[c].set_bypass(i, v)
        if rng.random() < 0.2:
            tch = int(rng.integers(-1, C))
        blk = np.ascontiguousarray(x[:, k * n:(k + 1) * n])
        ref = orc.controller_process(oc, ta, tch, sa, blk, fs)
        ma.process(list(blk) + list(ref[:C]) + [None] + list(ref[C:]), fs)
        if rng.random() < 0.5:
            got = gdg.controller_process(gc, tb, tch, sb, blk, fs, meter=mb)
        else:
            got = np.zeros((C + 2, n))
            t = gdg.controller_submit(gc, tb, tch, sb, blk, got, fs, meter=mb)
            gdg.controller_wait(t, lib)
        for c in range(C):
            assert common.rms(ref[c] - got[c]) <= common.TOL_RMS, (seed, k, c)
        # the master pair sums the channels (default positions: plain sum): its error is bounded by the channels' own
        scale = max(1.0, float(np.max(np.abs(ref[C:]))))
        bound = 2.0 * sum(float(np.max(np.abs(ref[c] - got[c]))) for c in range(C)) + 1e-11 * scale
        assert np.max(np.abs(ref[C:] - got[C:])) <= bound, (seed, k)
        if rng.random() < 0.3:
            ra, rb = ta.analyze(), tb.analyze()
            assert ra["note"] == rb["note"] and ra["cents"] == rb["cents"] and abs(ra["frequency"] - rb["frequency"]) <= 1e-6 * abs(ra["frequency"]), (seed, k, ra, rb)
    for c in range(2 * C + 3):
        assert ma.analyze(c) == mb.analyze(c), (seed, c)
        if c < C:                                                # input rows are bit-identical: so is the hold counter; on the
            assert ma.state(c) == pytest.approx(mb.state(c), rel=1e-10) and ma.state(c)[2] == mb.state(c)[2], (seed, c)   # output rows a 1e-16 difference can move the last ">= peak" sample


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("GDG_RANDOM_LONGIR", "6"))))
def test_random_long_filter_sessions(lib, seed):
    """Blocks of 8192 with composite power-amp filters of up to 65536 taps (the partitioned-convolution path: spectrum delay
    line, slot ring) while filters, levels and the order change between blocks and other units sit around the power amp."""
    rng = np.random.default_rng(17000 + seed)
    fs, n = int(rng.choice([96000, 192000])), 8192
    ob, pb = common.make_banks(lib, synthetic=[("long_a", 30000, fs), ("long_b", 65536, fs), ("mid", 9000, fs)])
    names = ["long_a", "long_b", "mid", "- NONE -"] + [str(v) for v in np.load(common.GOLD + "/ir_bank.npz")["names"][:3]]
    spec = [("compressor", {}, {}), ("power_amp", {}, {"filter_order": "65536", "filter_1": "long_a"}), ("delay", {"delay_time": 20}, {})]
    a, b = common.make_chain_pair(lib, spec, (ob, pb))
    x = common.synth_fast(seed, fs, 8 * n, seed)
    for k in range(8):
        r = rng.random()
        if r < 0.25:
            slot, name = int(rng.integers(1, 9)), str(rng.choice(names))
            a.set_discrete_value(1, "filter_%d" % slot, name); b.set_discrete_value(1, "filter_%d" % slot, name)
        elif r < 0.4:
            slot, lv = int(rng.integers(1, 9)), int(rng.integers(-30, 1))
            a.set_numeric_value(1, "level_%d" % slot, lv); b.set_numeric_value(1, "level_%d" % slot, lv)
        elif r < 0.55:
            order = str(rng.choice(["4096", "16384", "32768", "65536"]))
            a.set_discrete_value(1, "filter_order", order); b.set_discrete_value(1, "filter_order", order)
        elif r < 0.65:
            v = bool(rng.integers(2)); a.set_bypass(1, v); b.set_bypass(1, v)
        blk = x[k * n:(k + 1) * n]
        ya, yb = a.process(blk, fs), b.process(blk, fs)
        assert np.all(np.isfinite(yb)) and common.rms(ya - yb) <= common.TOL_RMS, (seed, k, common.rms(ya - yb))
