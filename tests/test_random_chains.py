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
    # discontinuous units (gate open/close, octaver polarity flips with hysteresis) amplify a 1e-18 difference of their
    # input into a different decision - between Go and any restatement too.  Like CHAIN12 the random chains
    # keep them where their input is bit-identical: gates first, then at most one octaver, then everything else.
    gates = [u for u in spec if u[0] == "noise_gate"]
    octs = [u for u in spec if u[0] == "octaver"][:1]
    spec = gates + octs + [u for u in spec if u[0] not in ("noise_gate", "octaver")]
    a, b = common.make_chain_pair(lib, spec, banks)
    blocks = 3 if n >= 2048 else 6
    x = common.synth_fast(seed, fs, blocks * n, seed) * float(rng.choice([0.1, 1.0, 1.7]))
    common.compare_stream(a, b, x, n, fs, label="seed %d fs %d n %d %s" % (seed, fs, n, [s[0] for s in spec]))


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


CONTINUOUS = [u for u in gdg.UNIT_TYPES if u not in ("noise_gate", "octaver")]


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
