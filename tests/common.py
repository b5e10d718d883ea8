"""Shared helpers of the parity tests: backends, synthetic inputs, canonical chains.

Two backends expose the SAME C ABI (include/gdg.h):
  "cuda" - go-dsp-guitar_b200/libgdg.so, the product (tests marked gpu);
  "emu"  - tests/emu/libgdg_emu.so, the same host runtime + the same stage source
           compiled as host code (tests only; checks the host logic and the parallel
           reformulation on CPU).
The checker is always the CPU oracle (oracle/pyoracle.py).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import gdg  # noqa: E402  (alias of the go-dsp-guitar_b200 package)
import pyoracle as orc  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
TOL_RMS = 1e-9          # BASELINE.json north_star: "within 1e-9 RMS float64"

_libs = {}


def backend_lib(kind):
    if kind not in _libs:
        if kind == "emu":
            sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
            import build_emu
            _libs[kind] = gdg.load(build_emu.build(), _test_backend=True)
        elif kind == "cuda":
            lib = gdg.load()                       # fails loudly if libgdg.so is missing
            gdg.init(-1, lib)                      # fails loudly without a B200
            _libs[kind] = lib
        else:
            raise ValueError(kind)
    return _libs[kind]


def rms(a):
    a = np.asarray(a, dtype=np.float64)
    return float(np.sqrt(np.mean(a * a))) if a.size else 0.0


# ---- synthetic input: two partials + reference-LCG noise, gated envelope ------------
def lcg_uniform(seed, n):
    """1 - 2*NextFloat() of random.CreatePRNG(seed) (random/random.go:23-55), vectorised exactly."""
    m = 2147483647
    x = ((64979 * (seed & 0xFFFFFFFFFFFFFFFF)) + 83) % (1 << 64) % m
    out = np.empty(n)
    xs = np.empty(n, dtype=np.int64)
    for i in range(n):
        x = (16807 * x) % m
        xs[i] = x
    out = 1.0 - 2.0 * (xs.astype(np.float64) / float(m - 1))
    return out


def synth_channel(c, fs, n_total, noise=None):
    t = np.arange(n_total, dtype=np.float64)
    f = 82.4069 * 2.0 ** ((c % 36) / 12.0)
    u = lcg_uniform(1337 + c, n_total) if noise is None else noise
    x = 0.4 * np.sin(2 * np.pi * f * t / fs) + 0.2 * np.sin(2 * np.pi * 2.01 * f * t / fs) + 0.05 * u
    env = np.minimum(1.0, t / (0.01 * fs))
    period = int(1.25 * fs)
    gate = ((t.astype(np.int64) % period) < int(1.0 * fs)).astype(np.float64)
    return x * env * gate


def synth_fast(c, fs, n_total, seed=0):
    """Same shape as synth_channel but numpy noise (for long / many-channel runs)."""
    rng = np.random.default_rng(1000 * seed + c)
    return synth_channel(c, fs, n_total, noise=rng.uniform(-1.0, 1.0, n_total))


# ---- impulse-response banks ---------------------------------------------------------------------------
def fill_bank_from_fixture(bank, decode):
    z = np.load(os.path.join(GOLD, "ir_bank.npz"))
    for k, (name, comp) in enumerate(zip(z["names"], z["compensation"])):
        bank.add(str(name), int(comp), decode(z["pcm_%02d" % k]), int(z["sample_rate"]))


def synthetic_ir(L, seed=7):
    """h[k] = (1 - 2 LCG_k) e^{-6k/L}, ||h||_2 = 0.1."""
    u = lcg_uniform(seed, L)
    h = u * np.exp(-6.0 * np.arange(L) / L)
    return h * (0.1 / np.sqrt(np.sum(h * h)))


def make_banks(lib, with_fixture=True, synthetic=()):
    """(oracle bank, product bank) with identical content."""
    ob, pb = orc.IRBank(), gdg.IRBank(lib)
    if with_fixture:
        fill_bank_from_fixture(ob, orc.pcm16_decode)
        fill_bank_from_fixture(pb, orc.pcm16_decode)      # decode is test plumbing; the product's own decoder is tested via import
    for name, L, rate in synthetic:
        h = synthetic_ir(L)
        ob.add_raw(name, 0, h, rate)
        pb.add_raw(name, 0, h, rate)
    return ob, pb


# ---- canonical 12-stage chain (SURVEY.md 8d CHAIN12) ----------------------------------------------------
CHAIN12 = [
    ("noise_gate", {}, {}),
    ("compressor", {"gain_limit": 30, "target_level": -20}, {"follow": "level"}),
    ("excess", {"gain": 6}, {"oversampling": "4"}),
    ("fuzz", {}, {"oversampling": "4"}),
    ("overdrive", {"gain": 20, "drive": 100}, {"valve": "ECC83 (12AX7)", "oversampling": "4"}),
    ("distortion", {"gain": 10, "level": -6}, {"oversampling": "4"}),
    ("tone_stack", {"low": 0, "middle": -2, "presence": -5, "high": -5}, {}),
    ("chorus", {"depth": 100, "speed": 30}, {}),
    ("delay", {"delay_time": 200, "feedback": -5, "level": -5}, {}),
    ("reverb", {"mix": 50}, {}),
    ("power_amp", {"level_1": 0}, {"filter_order": "8192", "filter_1": None}),
    ("cabinet", {}, {}),
]


def build_chain(chain, spec, ir_name=None):
    """Apply a [(unit, numeric, discrete)] spec through the signal.Chain API; every unit un-bypassed."""
    for unit, numeric, discrete in spec:
        i = chain.append_unit(unit)
        for k, v in discrete.items():
            if v is None:
                v = ir_name
            if v is not None:
                chain.set_discrete_value(i, k, v)
        for k, v in numeric.items():
            chain.set_numeric_value(i, k, v)
        chain.set_bypass(i, False)
    return chain


def make_chain_pair(lib, spec, banks=None, ir_name=None):
    ob, pb = banks if banks is not None else (None, None)
    a = build_chain(orc.Chain(ob), spec, ir_name)
    b = build_chain(gdg.Chain(pb, lib=lib), spec, ir_name)
    return a, b


def compare_stream(a, b, x, n, fs, tol=TOL_RMS, label="", rel_tol=1e-8, expect_signal=True):
    """Feed consecutive blocks of x (len multiple of n) through oracle chain a and product chain b.
    Per block: rms(err) <= tol (absolute, BASELINE.json's 1e-9) AND rms(err) <= rel_tol * rms(ref) wherever the reference
    block is above the noise floor, so that a quiet signal cannot hide a relative error; over the stream the reference must
    carry signal (rms > 0) unless the caller says otherwise, so a muted / gated output can never pass trivially."""
    worst = 0.0
    energy = 0.0
    for k in range(0, len(x), n):
        blk = x[k:k + n]
        ya = a.process(blk, fs)
        yb = b.process(blk, fs)
        assert np.all(np.isfinite(yb)), "%s: non-finite output in block %d" % (label, k // n)
        e, r = rms(ya - yb), rms(ya)
        energy = max(energy, r)
        worst = max(worst, e)
        assert e <= tol, "%s: block %d rms error %.3e > %.1e (peak ref %.3f)" % (label, k // n, e, tol, np.abs(ya).max())
        if r > 1e-6:
            assert e <= rel_tol * r, "%s: block %d relative rms error %.3e > %.1e (rms ref %.3e)" % (label, k // n, e / r, rel_tol, r)
    if expect_signal and len(x):
        assert energy > 0.0, "%s: the reference output is silent - the comparison proves nothing" % label
    return worst
