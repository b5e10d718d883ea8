"""Host-side checks that need no GPU: the C ABI surface, error conventions, IR import."""
import ctypes
import json
import os
import re
import struct
import subprocess

import numpy as np
import pytest

import common
from common import gdg, orc

ROOT = common.ROOT


def header_symbols():
    src = open(os.path.join(ROOT, "include", "gdg.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gdg_[a-z0-9_]+)\s*\(", src)))


def test_header_and_ctypes_table_agree():
    assert header_symbols() == sorted(gdg.ABI)


def test_product_library_loads_and_exports_every_declared_symbol():
    """libgdg.so (the CUDA build) must dlopen without a GPU and export everything gdg.h declares.
    No compute call is made here."""
    so = os.path.join(ROOT, "go-dsp-guitar_b200", "libgdg.so")
    if not os.path.exists(so):
        import importlib
        importlib.import_module("go-dsp-guitar_b200.build").build()
    lib = ctypes.CDLL(so)
    for name in header_symbols():
        assert hasattr(lib, name), name
    nm = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (gdg_[a-z0-9_]+)", nm))
    assert set(header_symbols()) <= exported
    assert gdg.load(so).gdg_backend_name().decode() == "cuda:sm_100a"
    # the oracle is not linked into the product
    assert "orc_" not in nm


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = gdg.load()
    assert lib.gdg_init(-1) != 0
    assert b"no CPU fallback" in lib.gdg_last_error()
    with pytest.raises(gdg.GdgError):
        gdg.Chain(lib=lib)


def test_unit_type_names_and_parameter_schema():
    lib = common.backend_lib("emu")
    assert lib.gdg_unit_type_count() == 21
    assert [lib.gdg_unit_type_name(i).decode() for i in range(21)] == gdg.UNIT_TYPES
    c = gdg.Chain(lib=lib)
    o = orc.Chain()
    for t in range(21):
        i = c.append_unit(t); o.append_unit(t)
        assert c.unit_type(i) == t
        for p in c.parameters(i):
            if p["Type"] == 2:
                assert o.get_numeric_value(i, p["Name"]) == p["NumericValue"]
                assert p["Minimum"] <= p["NumericValue"] <= p["Maximum"]
            else:
                assert o.get_discrete_value(i, p["Name"]) == p["DiscreteValues"][p["DiscreteValueIndex"]]
    od = {p["Name"]: p for p in c.parameters(9)}
    assert od["valve"]["DiscreteValues"] == ["ECC82 (12AU7)", "ECC83 (12AX7)"] and od["gain"]["PhysicalUnit"] == "dB"


def test_error_strings_match_the_reference():
    """effects/effects.go:165-312 and signal/signal.go:59-337 message formats."""
    lib = common.backend_lib("emu")
    c = gdg.Chain(lib=lib)
    o = orc.Chain()
    i = c.append_unit("overdrive"); o.append_unit("overdrive")

    def both(fn_name, *a):
        msgs = []
        for ch, exc in ((c, gdg.GdgError), (o, orc.OracleError)):
            with pytest.raises(exc) as e:
                getattr(ch, fn_name)(*a)
            msgs.append(str(e.value))
        assert msgs[0] == msgs[1], msgs
        return msgs[0]
    assert both("set_numeric_value", i, "gain", 31) == "Failed to set numeric value: Parameter 'gain' must be between '-30' and '30' - got '31'."
    assert both("set_numeric_value", i, "nope", 1) == "Failed to set numeric value: Could not find parameter with name 'nope'."
    assert both("set_numeric_value", i, "valve", 1) == "Failed to set numeric value: Parameter 'valve' is not numeric."
    assert both("set_discrete_value", i, "valve", "EL34") == "Failed to set discrete value: Value 'EL34' is not valid for parameter 'valve'."
    assert both("set_discrete_value", i, "gain", "x") == "Failed to set discrete value: Parameter 'gain' is not discrete."
    assert both("get_discrete_value", i, "gain") == "Failed to get discrete value: Parameter 'gain' is not discrete."
    assert both("set_numeric_value", 7, "gain", 1) == "Cannot set numeric value: No unit 7."
    assert both("set_bypass", 3, True) == "Cannot enable bypass: No unit 3."
    assert both("set_bypass", 3, False) == "Cannot disable bypass: No unit 3."
    assert both("remove_unit", -1) == "Cannot remove unit -1."
    assert both("move_up", 0) == "Cannot move unit 0 up."
    assert both("move_down", 0) == "Cannot move unit 0 down."
    assert both("unit_type", 5) == "Cannot get unit type: No unit 5."
    assert both("append_unit", 21) == "Failed to create effects unit."
    assert c.get_numeric_value(i, "gain") == 0                     # rejected value left the parameter unchanged
    # chain editing keeps ids dense like the reference's slice operations
    j = c.append_unit("delay"); k = c.append_unit("reverb")
    c.move_up(k); assert [c.unit_type(q) for q in range(3)] == [9, 18, 17]
    c.remove_unit(0); assert c.length() == 2 and c.unit_type(0) == 18


def test_block_length_contract():
    lib = common.backend_lib("emu")
    c = gdg.Chain(lib=lib)
    for n in (0, 3, 100, 16384):
        with pytest.raises(gdg.GdgError, match="power of two"):
            c.process(np.zeros(n), 48000)


def _write_wav16(path, pcm, rate, channels=1):
    data = np.asarray(pcm, dtype="<i2").tobytes()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, 1, channels, rate, rate * 2 * channels, 2 * channels, 16))
        f.write(b"data" + struct.pack("<I", len(data)) + data)


def test_irbank_import_json_and_wav(tmp_path, monkeypatch):
    """filter.Import (filter/filter.go:704-802) through the product's own JSON/WAV readers vs the oracle."""
    lib = common.backend_lib("emu")
    z = np.load(os.path.join(common.GOLD, "ir_bank.npz"))
    monkeypatch.chdir(tmp_path)
    os.makedirs("ir")
    index = []
    for k in (3, 9, 23):
        _write_wav16("ir/r%d.wav" % k, z["pcm_%02d" % k], 96000)
        index.append({"Name": str(z["names"][k]), "Path": "ir/r%d.wav" % k, "Compensation": int(z["compensation"][k])})
    _write_wav16("ir/stereo.wav", np.zeros(64), 96000, channels=2)
    index.append({"Name": "stereo (skipped)", "Path": "ir/stereo.wav", "Compensation": 0})
    index.append({"Name": "missing (skipped)", "Path": "ir/nope.wav", "Compensation": 0})
    json.dump(index, open("index.json", "w"))
    pb = gdg.IRBank.import_file("index.json", lib)
    ob = orc.IRBank()
    for k in (3, 9, 23):
        ob.add(str(z["names"][k]), int(z["compensation"][k]), orc.pcm16_decode(z["pcm_%02d" % k]), 96000)
    assert pb.names() == ob.names() == [str(z["names"][k]) for k in (3, 9, 23)]
    for fs, name in ((44100, pb.names()[0]), (192000, pb.names()[2])):
        spec = [("power_amp", {}, {"filter_order": "4096", "filter_1": name})]
        a, b = common.make_chain_pair(lib, spec, (ob, pb))
        x = common.synth_fast(1, fs, 3 * 2048)
        common.compare_stream(a, b, x, 2048, fs, label="import")
    with pytest.raises(gdg.GdgError, match="Failed to read descriptor file"):
        gdg.IRBank.import_file("nope.json", lib)


@pytest.mark.parametrize("fmt", [0, 1, 2, 3, 4, 5])
def test_irbank_import_every_wave_format(tmp_path, monkeypatch, fmt):
    """filter.Import reads its files through wave.FromBuffer: all six payload formats of wave.go:790-838 (user impulse
    responses in 24-bit PCM or float32 are common).  The imported response must equal the oracle's decode of the same bytes."""
    lib = common.backend_lib("emu")
    monkeypatch.chdir(tmp_path)
    rng = np.random.default_rng(fmt)
    h = rng.uniform(-0.9, 0.9, 257) * np.exp(-np.arange(257) / 40.0)
    payload = orc.wave_encode(fmt, h).tobytes()
    bits = [8, 16, 24, 32, 32, 64][fmt]
    tag = 1 if fmt < 4 else 3
    with open("ir.wav", "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(payload)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, tag, 1, 48000, 48000 * bits // 8, bits // 8, bits))
        f.write(b"data" + struct.pack("<I", len(payload)) + payload)
    json.dump([{"Name": "user ir", "Path": "ir.wav", "Compensation": -2}], open("index.json", "w"))
    pb = gdg.IRBank.import_file("index.json", lib)
    assert pb.names() == ["user ir"]
    ob = orc.IRBank()
    ob.add("user ir", -2, orc.wave_decode(fmt, np.frombuffer(payload, dtype=np.uint8)), 48000)
    a, b = common.make_chain_pair(lib, [("power_amp", {}, {"filter_1": "user ir"})], (ob, pb))
    x = common.synth_fast(2, 48000, 2 * 1024)
    common.compare_stream(a, b, x, 1024, 48000, label="import fmt %d" % fmt)


def test_parameter_table_matches_reference_source():
    if not os.path.isdir("/root/reference/effects"):
        pytest.skip("reference tree not present (GPU box)")
    r = subprocess.run(["python", os.path.join(ROOT, "tools", "check_unit_params.py")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
