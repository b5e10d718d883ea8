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


# ---- synthetic input, canonical chain: defined once in the package (go-dsp-guitar_b200/presets.py) -----------
import importlib  # noqa: E402
_presets = importlib.import_module("go-dsp-guitar_b200.presets")
lcg_uniform, synth_channel, synth_fast = _presets.lcg_uniform, _presets.synth_channel, _presets.synth_fast
synthetic_ir, CHAIN12, build_chain = _presets.synthetic_ir, _presets.CHAIN12, _presets.build_chain


# ---- impulse-response banks ---------------------------------------------------------------------------
def fill_bank_from_fixture(bank, decode):
    z = np.load(os.path.join(GOLD, "ir_bank.npz"))
    for k, (name, comp) in enumerate(zip(z["names"], z["compensation"])):
        bank.add(str(name), int(comp), decode(z["pcm_%02d" % k]), int(z["sample_rate"]))


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
