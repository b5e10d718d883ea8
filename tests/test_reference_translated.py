"""Parity against golden vectors produced by EXECUTING the reference's own Go sources.

There is no Go toolchain here or on the GPU box, so tests/golden/go2py/ translates the reference's .go files mechanically
(statement by statement, no knowledge of the DSP) into Python, checks the translation against the reference's OWN unit-test
vectors (fft, oversampling, resample, random, circular, level, wave codecs, tuner: 72 vectors) and then runs every unit, the chain driver,
the power amp's IR algebra, the spatializer and the level meter to produce tests/golden/translated_reference_vectors.npz
(generator: tests/golden/go2py/make_reference_vectors.py; it needs /root/reference, the committed vectors do not).

This pins the rows the reference itself holds no test for (a1, a2, a8 - a20): the C oracle must reproduce the
translated reference at <= 1e-12 (relative to the peak), and the product - the host emulation on CPU, libgdg.so on the B200 -
at the 1e-9 RMS of BASELINE.json.  Nobody's reading of the Go code is in this loop."""
import json
import os

import numpy as np
import pytest

import common
from common import gdg, orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
M = json.load(open(os.path.join(GOLD, "translated_reference_vectors.json")))
Z = np.load(os.path.join(GOLD, "translated_reference_vectors.npz"))
CASES = M["cases"]
UNIT_NAMES = M["unit_names"]
BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]
TOL_ORACLE = 1e-12                    # oracle vs translated reference: libm vs Go's math (<= 1-2 ulp per call), nothing else
GROUPS = sorted(set((c["kind"], c["spec"][0][0] if (c["kind"] == "unit" and c["spec"]) else c["kind"]) for c in CASES))


@pytest.fixture(scope="module")
def oracle_bank():
    b = orc.IRBank()
    common.fill_bank_from_fixture(b, orc.pcm16_decode)
    return b


def apply_edit(chain, op):
    name, args = op[0], op[1:]
    if name == "AppendUnit":
        return chain.append_unit(UNIT_NAMES[args[0]])
    return {"MoveUp": chain.move_up, "MoveDown": chain.move_down, "RemoveUnit": chain.remove_unit, "SetBypass": chain.set_bypass,
            "SetNumericValue": chain.set_numeric_value, "SetDiscreteValue": chain.set_discrete_value}[name](*args)


def run_case(chain, case):
    x = case["amp"] * Z[case["input"]]
    ys, pos = [], 0
    for k, n in enumerate(case["ns"]):
        for op in case["edits"].get(str(k), ()):
            apply_edit(chain, op)
        ys.append(chain.process(np.ascontiguousarray(x[pos:pos + n]), case["fss"][k]))
        pos += n
    return np.concatenate(ys)


def spec_of(case):
    return [(u, nu, di) for u, nu, di in case["spec"]]


def test_manifest_is_complete():
    """every unit type, every discrete value of every unit and both chain kinds are in the vectors; the generator's
    self-check reproduced the reference's own test vectors first"""
    assert M["selfcheck_vectors"] >= 60
    units = set(c["spec"][0][0] for c in CASES if c["kind"] == "unit")
    assert units == set(UNIT_NAMES) - {"power_amp"} and len(UNIT_NAMES) == 21
    assert sum(c["kind"] == "power_amp" for c in CASES) >= 6 and sum(c["kind"] == "chain" for c in CASES) >= 5
    assert sum(1 for c in CASES if c["peak"] > 0) >= len(CASES) - 10         # silent: zero amplitudes, closed gates, reverb tail not reached, no filter
    for name in ("excess", "fuzz", "overdrive", "distortion"):
        seen = set(c["spec"][0][2].get("oversampling") for c in CASES if c["kind"] == "unit" and c["spec"][0][0] == name)
        assert {"- NONE -", "2", "4"} <= seen, (name, seen)


@pytest.mark.parametrize("kind,group", GROUPS)
def test_oracle_reproduces_translated_reference(kind, group, oracle_bank):
    n_cases = 0
    for case in CASES:
        if case["kind"] != kind or (kind == "unit" and case["spec"][0][0] != group):
            continue
        chain = common.build_chain(orc.Chain(oracle_bank), spec_of(case), case["ir"])
        y = run_case(chain, case)
        ref = Z[case["id"]]
        err = float(np.max(np.abs(y - ref))) if ref.size else 0.0
        assert np.all(np.isfinite(y))
        assert err <= TOL_ORACLE * max(1.0, float(np.max(np.abs(ref)))), "%s (%s): max |oracle - reference| = %.3e" % (case["id"], case["label"], err)
        n_cases += 1
    assert n_cases > 0


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("kind,group", GROUPS)
def test_product_reproduces_translated_reference(kind, group, backend):
    """the C ABI (host emulation of the kernel source on CPU, libgdg.so on the B200) against the reference's own arithmetic:
    1e-9 RMS per block (BASELINE.json), and relative to the block where it carries signal"""
    lib = common.backend_lib(backend)
    _ob, bank = common.make_banks(lib)
    n_cases = 0
    for case in CASES:
        if case["kind"] != kind or (kind == "unit" and case["spec"][0][0] != group):
            continue
        chain = common.build_chain(gdg.Chain(bank, lib=lib), spec_of(case), case["ir"])
        y = run_case(chain, case)
        ref = Z[case["id"]]
        pos = 0
        for k, n in enumerate(case["ns"]):
            e = common.rms(y[pos:pos + n] - ref[pos:pos + n])
            r = common.rms(ref[pos:pos + n])
            assert e <= common.TOL_RMS, "%s (%s) block %d: rms error %.3e" % (case["id"], case["label"], k, e)
            if r > 1e-6:
                assert e <= 1e-8 * r, "%s (%s) block %d: relative rms error %.3e" % (case["id"], case["label"], k, e / r)
            pos += n
        n_cases += 1
    assert n_cases > 0


# ---- spatializer.Process ---------------------------------------------------------------------------------------------
def _spat_setup(sp, geo, fs):
    sp.set_sample_rate(fs)
    for c, (az, dist, lvl) in enumerate(geo):
        sp.set_azimuth(c, az); sp.set_distance(c, dist); sp.set_level(c, lvl)


def test_oracle_spatializer_reproduces_translated_reference():
    s = M["spatializer"]
    sp = orc.Spatializer(s["channels"])
    _spat_setup(sp, s["geometry"], s["fs"])
    x, aux, ref = Z["spat_in"], Z["spat_aux"], Z["spat_out"]
    n = s["n"]
    for k in range(s["blocks"]):
        y = sp.process(x[:, k * n:(k + 1) * n], aux[k * n:(k + 1) * n] if k % 2 == 0 else None)
        assert np.max(np.abs(np.asarray(y) - ref[:, k * n:(k + 1) * n])) <= 1e-13 * max(1.0, np.max(np.abs(ref))), k
    assert np.max(np.abs(ref)) > 0.1


@pytest.mark.parametrize("backend", BACKENDS)
def test_product_spatializer_reproduces_translated_reference(backend):
    lib = common.backend_lib(backend)
    s = M["spatializer"]
    sp = gdg.Spatializer(s["channels"], lib=lib)
    _spat_setup(sp, s["geometry"], s["fs"])
    x, aux, ref = Z["spat_in"], Z["spat_aux"], Z["spat_out"]
    n = s["n"]
    for k in range(s["blocks"]):
        y = sp.process(np.ascontiguousarray(x[:, k * n:(k + 1) * n]), np.ascontiguousarray(aux[k * n:(k + 1) * n]) if k % 2 == 0 else None)
        assert np.max(np.abs(np.asarray(y) - ref[:, k * n:(k + 1) * n])) <= 1e-12 * max(1.0, np.max(np.abs(ref))), k


def test_spatializer_at_other_rates_like_the_executed_reference():
    """spatializer.SetSampleRate never stores the rate (spatializer.go:418-431), so the inter-aural delays stay those of 96 kHz:
    the executed reference at 192 kHz and 44.1 kHz against the oracle and the emulated product (CPU)"""
    s = M["spatializer"]
    x, n = Z["spat_in"], s["n"]
    lib = common.backend_lib("emu")
    for rate in s["other_rates"]:
        ref = Z["spat_out_%d" % rate]
        for sp in (orc.Spatializer(s["channels"]), gdg.Spatializer(s["channels"], lib=lib)):
            _spat_setup(sp, s["geometry"], rate)
            for k in range(s["blocks"]):
                y = sp.process(np.ascontiguousarray(x[:, k * n:(k + 1) * n]), None)
                assert np.max(np.abs(np.asarray(y) - ref[:, k * n:(k + 1) * n])) <= 1e-12 * max(1.0, np.max(np.abs(ref))), (rate, k)


# ---- filter.Import / Reduce / Normalize / Multiply ---------------------------------------------------------------------
def _fixture_ir(name):
    z = np.load(os.path.join(GOLD, "ir_bank.npz"))
    k = list(z["names"]).index(name)
    return orc.pcm16_decode(z["pcm_%02d" % k]), int(z["sample_rate"]), int(z["compensation"][k])


def test_oracle_import_and_algebra_reproduce_translated_reference():
    for info in M["ir"]:
        pcm, rate, _comp = _fixture_ir(info["name"])
        got = orc.resample_time(pcm, rate, info["rate"])                 # filter.go:766 resample.Time per supported rate
        ref = Z[info["key"]]
        assert len(got) == len(ref), info
        assert np.max(np.abs(got - ref)) <= 1e-13, info
    names = M["ir_names"]
    p1, r1, c1 = _fixture_ir(names[0])
    p2, r2, c2 = _fixture_ir(names[1])
    f1 = orc.Filter(orc.resample_time(p1, r1, 96000), 96000); f1.set_compensation(10.0 ** (0.05 * c1))
    f2 = orc.Filter(orc.resample_time(p2, r2, 96000), 96000); f2.set_compensation(10.0 ** (0.05 * c2))
    for got, key in ((f2.reduce(1024).coefficients(), "alg_reduce_1024"), (f1.normalize().coefficients(), "alg_normalize"),
                     (f1.multiply(0.5).coefficients(), "alg_multiply")):
        ref = Z[key]
        assert len(got) == len(ref), key
        assert np.max(np.abs(np.asarray(got) - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref))), key
    a, b = np.asarray(f1.normalize().coefficients()), np.asarray(f2.normalize().coefficients())
    s = np.zeros(max(len(a), len(b))); s[:len(a)] += a; s[:len(b)] += b
    assert len(s) == len(Z["alg_add"]) and np.max(np.abs(s - Z["alg_add"])) <= 1e-12 * max(1.0, np.max(np.abs(s)))


@pytest.mark.parametrize("backend", BACKENDS)
def test_product_compiled_filter_reproduces_translated_reference(backend):
    """the power amp's compiled coefficients (device-side Import resampling -> Reduce -> Normalize -> Multiply -> Add) against
    the translated filter algebra: one response at level 0 with filter_order 1024 = Normalize(Reduce(1024))"""
    lib = common.backend_lib(backend)
    _ob, bank = common.make_banks(lib)
    names = M["ir_names"]
    ch = common.build_chain(gdg.Chain(bank, lib=lib), [("power_amp", {"level_1": 0}, {"filter_order": "1024", "filter_1": names[1]})])
    ch.process(np.zeros(256), 96000)                                      # compiles for this rate
    got = np.asarray(ch.filter_coefficients(0))
    p2, r2, c2 = _fixture_ir(names[1])
    f2 = orc.Filter(Z["ir_1_96000"], 96000); f2.set_compensation(10.0 ** (0.05 * c2))
    want = np.asarray(orc.Filter(Z["alg_reduce_1024"], 96000).coefficients())
    assert len(got) == len(want)
    fr = orc.Filter(Z["alg_reduce_1024"], 96000); fr.set_compensation(10.0 ** (0.05 * c2))
    want = np.asarray(fr.normalize().coefficients())
    assert np.max(np.abs(got - want)) <= 1e-11 * max(1.0, np.max(np.abs(want)))


# ---- level.Meter ---------------------------------------------------------------------------------------------------------
def test_oracle_meter_reproduces_translated_reference():
    s = M["level_meter"]
    m = orc.Meter(2)
    m.set_enabled(True)
    x, n = Z["meter_in"], s["n"]
    for k in range(s["blocks"]):
        m.process([x[c, k * n:(k + 1) * n] for c in range(2)], s["fs"])
        assert [list(m.analyze(c)) for c in range(2)] == s["readouts"][k], k
    assert len(set(tuple(r[0]) for r in s["readouts"])) > 3                    # the read-outs move: decay and peak hold are exercised


@pytest.mark.parametrize("backend", BACKENDS)
def test_product_meter_reproduces_translated_reference(backend):
    lib = common.backend_lib(backend)
    s = M["level_meter"]
    m = gdg.Meter(2, lib=lib)
    m.set_enabled(True)
    x, n = Z["meter_in"], s["n"]
    for k in range(s["blocks"]):
        m.process([np.ascontiguousarray(x[c, k * n:(k + 1) * n]) for c in range(2)], s["fs"])
        assert [list(m.analyze(c)) for c in range(2)] == s["readouts"][k], k


def test_reference_panics_where_the_documented_difference_says():
    """flanger / phaser at 22.05 / 44.1 / 88.2 kHz with depth 100 read one sample before their history.  The
    generator ran the translated reference there: it panics ('index out of range [-1]'), and not at 48 kHz."""
    got = {(p["unit"], p["fs"]): p["panic"] for p in M["reference_panics"]}
    for key in (("flanger", 44100), ("phaser", 22050), ("flanger", 88200)):
        assert got[key] is not None and "index out of range [-1]" in got[key]["message"], key
    assert got[("flanger", 48000)] is None


def test_config1_pcm16_bytes_equal_the_executed_reference():
    """BASELINE.json config 1 in miniature (1-ch overdrive -> tone_stack, 48 kHz, PCM16 in and out, blocks of 8192 with the last one
    zero-padded): the output payload of the executed reference (wave.go's conversions + signal.Chain), byte for byte, from the
    oracle and from the product's fused codecs (host emulation; the same path runs on the GPU in tests/test_parity.py)"""
    c = M["config1"]
    n, nb = c["block"], c["blocks"]
    raw = bytes(Z["cfg1_pcm_in"].tobytes()) + b"\x00" * (2 * (n * nb - c["samples"]))
    want = Z["cfg1_pcm_out"].tobytes()
    spec = [(u, nu, di) for u, nu, di in c["spec"]]
    oc = common.build_chain(orc.Chain(), spec)
    got = b"".join(orc.wave_encode("PCM16", oc.process(orc.wave_decode("PCM16", raw[2 * n * k:2 * n * (k + 1)]), c["fs"])).tobytes() for k in range(nb))
    assert got == want
    lib = common.backend_lib("emu")
    pc = common.build_chain(gdg.Chain(lib=lib), spec)
    got = b"".join(gdg.process_batch_encoded([pc], [raw[2 * n * k:2 * n * (k + 1)]], "PCM16", "PCM16", n, c["fs"])[0] for k in range(nb))
    assert got == want


# ---- parameter tables, error strings (effects.Unit / signal.Chain API surface) ------------------------
def _three_response_banks(lib):
    """the responses the generator imported (manifest ir_names), in index.json order like filter.Import keeps them"""
    z = np.load(os.path.join(GOLD, "ir_bank.npz"))
    ob, pb = orc.IRBank(), gdg.IRBank(lib)
    for k, name in enumerate(z["names"]):
        if str(name) in M["ir_names"]:
            for b in (ob, pb):
                b.add(str(name), int(z["compensation"][k]), orc.pcm16_decode(z["pcm_%02d" % k]), int(z["sample_rate"]))
    return ob, pb


def test_parameter_tables_equal_the_executed_reference():
    """names, types, physical units, ranges, defaults and discrete value lists of all 21 units - the power amp's after
    effects.PreparePowerAmp (filter_1..8 over the imported response names) - as the translated reference's Parameters() returns them"""
    lib = common.backend_lib("emu")
    ob, pb = _three_response_banks(lib)
    c, o = gdg.Chain(pb, lib=lib), orc.Chain(ob)
    for t, want in enumerate(M["parameters"]):
        i = c.append_unit(t); o.append_unit(t)
        got = c.parameters(i)
        assert [p["Name"] for p in got] == [p["Name"] for p in want], UNIT_NAMES[t]
        for g, w in zip(got, want):
            for key in ("Type", "PhysicalUnit", "Minimum", "Maximum", "NumericValue", "DiscreteValueIndex"):
                assert g[key] == w[key], (UNIT_NAMES[t], w["Name"], key, g[key], w[key])
            assert list(g["DiscreteValues"]) == w["DiscreteValues"], (UNIT_NAMES[t], w["Name"])
            if w["Type"] == 2:
                assert o.get_numeric_value(i, w["Name"]) == w["NumericValue"]
            elif w["Type"] == 1:
                assert o.get_discrete_value(i, w["Name"]) == w["DiscreteValues"][w["DiscreteValueIndex"]]
        assert c.get_bypass(i) is True and c.unit_type(i) == t


def test_error_strings_equal_the_executed_reference():
    lib = common.backend_lib("emu")
    E = M["errors"]
    for ch, exc in ((gdg.Chain(lib=lib), gdg.GdgError), (orc.Chain(), orc.OracleError)):
        i = ch.append_unit("overdrive")

        def fails(fn, *a):
            with pytest.raises(exc) as e:
                getattr(ch, fn)(*a)
            return str(e.value)
        assert fails("set_numeric_value", i, "gain", 31) == E["set_numeric_value out of range"]
        assert fails("set_numeric_value", i, "nope", 1) == E["set_numeric_value unknown"]
        assert fails("set_numeric_value", i, "valve", 1) == E["set_numeric_value on discrete"]
        assert fails("set_discrete_value", i, "valve", "EL34") == E["set_discrete_value bad value"]
        assert fails("set_discrete_value", i, "gain", "x") == E["set_discrete_value on numeric"]
        assert fails("get_discrete_value", i, "gain") == E["get_discrete_value on numeric"]
        assert fails("get_numeric_value", i, "nope") == E["get_numeric_value unknown"]
        assert fails("set_numeric_value", 7, "gain", 1) == E["set_numeric_value no unit"]
        assert fails("set_bypass", 3, True) == E["set_bypass true no unit"]
        assert fails("set_bypass", 3, False) == E["set_bypass false no unit"]
        assert fails("get_bypass", 3) == E["get_bypass no unit"]
        assert fails("remove_unit", -1) == E["remove_unit -1"]
        assert fails("move_up", 0) == E["move_up 0"]
        assert fails("move_down", 0) == E["move_down 0"]
        assert fails("unit_type", 5) == E["unit_type no unit"]
        assert fails("append_unit", 21) == E["append_unit 21"]
        assert ch.get_numeric_value(i, "gain") == 0


# ---- tuner.Process / Analyze ---------------------------------------------------------------------------------------------
def _tuner_check(make, analyze):
    for t in M["tuner"]:
        n = t["n"]
        tt = np.arange(t["blocks"] * n) / t["fs"]                  # make_reference_vectors.tuner_signal
        f0 = t["f0"]
        x = 0.5 * np.sin(2 * np.pi * f0 * tt) + 0.2 * np.sin(2 * np.pi * 2 * f0 * tt + 0.4) + 0.1 * np.sin(2 * np.pi * 3 * f0 * tt + 1.1)
        tu = make()
        for k in range(len(x) // n):
            tu.process(np.ascontiguousarray(x[k * n:(k + 1) * n]), t["fs"])
        note, cents, freq = analyze(tu)
        assert (note, cents) == (t["note"], t["cents"]), (t["f0"], note, cents)
        assert abs(freq - t["frequency"]) <= 1e-9 * t["frequency"], (t["f0"], freq, t["frequency"])
        assert abs(freq - t["f0"]) < 0.02 * t["f0"]


def _tuner_fields(r):
    if isinstance(r, dict):
        return r["note"], int(r["cents"]), float(r["frequency"])
    if hasattr(r, "note"):
        nt = r.note.decode() if isinstance(r.note, bytes) else r.note
        return nt, int(r.cents), float(r.frequency)
    return r[0], int(r[1]), float(r[2])


def test_oracle_tuner_reproduces_translated_reference():
    _tuner_check(orc.Tuner, lambda tu: _tuner_fields(tu.analyze()))


@pytest.mark.parametrize("backend", BACKENDS)
def test_product_tuner_reproduces_translated_reference(backend):
    lib = common.backend_lib(backend)
    _tuner_check(lambda: gdg.Tuner(lib=lib), lambda tu: _tuner_fields(tu.analyze()))
