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
