"""Canonical workloads of the hot path: the 12-stage chain CHAIN12, the synthetic input signal and the
synthetic impulse response.  Pure numpy + the package's own Chain API: shared by bench.py, examples/ and tests/ (the
product arm of bench.py must not import anything from tests/ or oracle/)."""
import numpy as np

# units in reference menu order (effects/effects.go:543-565); discontinuous units first so that they see bit-identical input
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
    """Apply a [(unit, numeric, discrete)] spec through the signal.Chain API (AppendUnit / SetDiscreteValue /
    SetNumericValue / SetBypass(false)); a discrete value of None stands for `ir_name`."""
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


def lcg_uniform(seed, n):
    """1 - 2*NextFloat() of random.CreatePRNG(seed) (random/random.go:23-55)."""
    m = 2147483647
    x = ((64979 * (seed & 0xFFFFFFFFFFFFFFFF)) + 83) % (1 << 64) % m
    xs = np.empty(n, dtype=np.int64)
    for i in range(n):
        x = (16807 * x) % m
        xs[i] = x
    return 1.0 - 2.0 * (xs.astype(np.float64) / float(m - 1))


def synth_channel(c, fs, n_total, noise=None):
    """two partials of a per-channel fundamental + reference-LCG noise, 10 ms fade-in, gated 1 s on / 0.25 s off"""
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


def synthetic_ir(L, seed=7):
    """h[k] = (1 - 2 LCG_k) e^{-6k/L}, ||h||_2 = 0.1."""
    u = lcg_uniform(seed, L)
    h = u * np.exp(-6.0 * np.arange(L) / L)
    return h * (0.1 / np.sqrt(np.sum(h * h)))
