"""Patch files (persistence.Configuration, persistence/persistence.go:6-88): examples/chain12.patch.json is BASELINE config 4's
chain set in the format the reference's own "save patch" button writes (controller.go:1195-1370).  The SAME file configures the
product, the oracle and - where the reference tree is present - the executed reference (go2py), and they agree on the audio."""
import importlib
import json
import os
import sys

import numpy as np
import pytest

import common
from common import gdg, orc

patch = importlib.import_module("go-dsp-guitar_b200.patch")
FILE = os.path.join(common.ROOT, "examples", "chain12.patch.json")
IR = "Guitar: American Vintage (Center)"
BACKENDS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]
REFERENCE = "/root/reference"
have_reference = os.path.isdir(os.path.join(REFERENCE, "persistence"))


@pytest.fixture(params=BACKENDS)
def lib(request):
    return common.backend_lib(request.param)


def _product_set(lib, channels, spec=()):
    _ob, pb = common.make_banks(lib)
    chains = [common.build_chain(gdg.Chain(pb, lib=lib), list(spec), IR) for _ in range(channels)]
    return chains, gdg.Spatializer(channels, lib=lib)


def test_committed_file_is_what_persist_writes_for_config_4(lib):
    chains, spat = _product_set(lib, 8, common.CHAIN12)
    for c in range(8):
        spat.set_azimuth(c, -70.0 + 20.0 * c); spat.set_distance(c, 1.0 + 0.25 * c); spat.set_level(c, 1.0)
    cfg = patch.persist(chains, spat, frames_per_period=8192)
    assert cfg == json.load(open(FILE))
    units = cfg["Channels"][0]["Units"]
    assert [u["Type"] for u in units] == [u[0] for u in common.CHAIN12] and not any(u["Bypass"] for u in units)
    # every parameter of every unit is in the file, like the reference writes it (controller.go:1255-1289)
    assert sum(len(u["DiscreteParams"]) + len(u["NumericParams"]) for u in units) == sum(len(chains[0].parameters(i)) for i in range(12))


def test_restore_rebuilds_chains_and_positions_and_the_audio_follows(lib):
    fs, n = 48000, 512
    cfg = json.load(open(FILE))
    chains, spat = _product_set(lib, 8, [("tremolo", {}, {}), ("delay", {}, {})])       # something else first: restore empties the chains
    assert patch.restore(open(FILE).read(), chains, spat) == []
    again = patch.persist(chains, spat, frames_per_period=8192)
    assert again == cfg                                                                  # file -> chains -> file is the identity
    assert spat.get_azimuth(5) == 30.0 and spat.get_distance(5) == 2.25 and spat.get_level(5) == 1.0
    ob, _pb = common.make_banks(lib)
    ochains = [orc.Chain(ob) for _ in range(8)]
    ospat = orc.Spatializer(8)
    assert patch.restore(cfg, ochains, ospat) == []                                      # the oracle is configured from the same file
    spat.set_sample_rate(fs); ospat.set_sample_rate(fs)
    x = np.stack([common.synth_fast(c, fs, 4 * n) for c in range(8)])
    for k in range(4):
        blk = x[:, k * n:(k + 1) * n]
        ref = np.stack([ochains[c].process(blk[c], fs) for c in range(8)])
        rows = gdg.controller_process(chains, None, -1, spat, blk, fs)                  # C output rows + master L / R
        out, master = rows[:8], rows[8:]
        assert common.rms(ref - np.asarray(out)) <= common.TOL_RMS and common.rms(ref) > 0
        assert common.rms(ospat.process(ref) - np.asarray(master)) <= common.TOL_RMS


def test_restore_keeps_going_like_the_reference_and_reports_what_it_ignored(lib):
    chains, spat = _product_set(lib, 2)
    cfg = {"fileformat": {"TYPE": "patch", "Version": {"Major": 1, "Minor": 3}},              # encoding/json folds the case of keys
           "Channels": [{"Units": [{"Type": "overdrive", "Bypass": True, "NumericParams": [{"Key": "gain", "Value": 999}, {"Key": "level", "Value": -3}],
                                    "DiscreteParams": [{"Key": "valve", "Value": "EL34"}]},
                                   {"Type": "kazoo"}, {"Type": "cabinet", "Bypass": False}],
                         "Spatializer": {"Azimuth": 45.0, "Distance": 99.0, "Level": 0.5}}]}
    ignored = patch.restore(json.dumps(cfg), chains, spat)
    assert chains[0].length() == 2 and chains[1].length() == 0
    assert chains[0].get_bypass(0) and not chains[0].get_bypass(1)
    assert chains[0].get_numeric_value(0, "level") == -3 and chains[0].get_numeric_value(0, "gain") == 0          # out of range: unchanged
    assert spat.get_azimuth(0) == 45.0 and spat.get_level(0) == 0.5 and spat.get_distance(0) == 0.0                # [0, 10] only
    assert len(ignored) == 4 and any("kazoo" in s for s in ignored) and any("must be between" in s for s in ignored)
    for bad, msg in (({"FileFormat": {"Type": "capture"}}, "not a patch file"), ({"FileFormat": {"Type": "patch", "Version": {"Major": 2}}}, "Incompatible version"),
                     ("{", "unmarshalling")):
        with pytest.raises(patch.PatchError, match=msg):
            patch.restore(bad if isinstance(bad, str) else json.dumps(bad), chains, spat)
    with pytest.raises(patch.PatchError, match="contains 3 channels"):
        patch.restore({"FileFormat": {"Type": "patch", "Version": {"Major": 1}}, "Channels": [{}, {}, {}]}, chains, spat)


# ---- with the reference tree: its own structs and its own chains read the same file -------------------------------------
class _GoChain(object):
    """the executed reference's signal.Chain behind the snake_case Chain API patch.restore() drives"""

    def __init__(self, ch):
        self.ch = ch

    def length(self):
        return int(self.ch.Length())

    def remove_unit(self, i):
        assert self.ch.RemoveUnit(i) is None

    def append_unit(self, t):
        uid, err = self.ch.AppendUnit(t)
        assert err is None, err
        return uid

    def set_discrete_value(self, i, k, v):
        e = self.ch.SetDiscreteValue(i, k, v)
        if e is not None:
            raise ValueError(e.Error())

    def set_numeric_value(self, i, k, v):
        e = self.ch.SetNumericValue(i, k, v)
        if e is not None:
            raise ValueError(e.Error())

    def set_bypass(self, i, b):
        assert self.ch.SetBypass(i, b) is None


@pytest.mark.skipif(not have_reference, reason="needs the reference tree (build container)")
def test_file_matches_the_reference_structs_and_its_chains_play_it_like_the_oracle():
    sys.path.insert(0, os.path.join(common.ROOT, "tests", "golden", "go2py"))
    import make_reference_vectors as gen
    from goruntime import G
    ref = gen.Ref(REFERENCE)
    pers = ref.loader.load("persistence")
    cfg = json.load(open(FILE))

    def conforms(value, go, path):
        """every key of the file names a field of the translated Go struct (exact case, like the reference's own writer)"""
        fields = {k for k in vars(go) if not k.startswith("_")}
        assert set(value) == fields, (path, sorted(set(value) ^ fields))
    conforms(cfg, pers.Configuration(), "Configuration")
    conforms(cfg["FileFormat"], pers.FileFormat(), "FileFormat"); conforms(cfg["FileFormat"]["Version"], pers.Version(), "Version")
    conforms(cfg["Metronome"], pers.Metronome(), "Metronome")
    for ch in cfg["Channels"]:
        conforms(ch, pers.Channel(), "Channel"); conforms(ch["Spatializer"], pers.Spatializer(), "Spatializer")
        for u in ch["Units"]:
            conforms(u, pers.Unit(), "Unit")
            assert all(set(p) == {"Key", "Value"} and isinstance(p["Value"], str) for p in u["DiscreteParams"])
            assert all(set(p) == {"Key", "Value"} and isinstance(p["Value"], int) for p in u["NumericParams"])
    assert list(ref.unit_names) == patch.UNIT_TYPES
    # one channel of the file through the reference's own chain and spatializer, executed; the oracle from the same file
    one = dict(cfg, Channels=cfg["Channels"][2:3])
    go_chain, go_spat = _GoChain(ref.signal.CreateChain(ref.irs)), ref.spatializer.Create(1)

    class _GoSpat(object):
        def set_azimuth(self, c, v): assert go_spat.SetAzimuth(c, v) is None
        def set_distance(self, c, v): assert go_spat.SetDistance(c, v) is None
        def set_level(self, c, v): assert go_spat.SetLevel(c, v) is None
    assert patch.restore(one, [go_chain], _GoSpat()) == []
    ob, _pb = common.make_banks(common.backend_lib("emu"))
    ochain, ospat = orc.Chain(ob), orc.Spatializer(1)
    assert patch.restore(one, [ochain], ospat) == []
    fs, n = 48000, 256
    go_spat.SetSampleRate(fs); ospat.set_sample_rate(fs)
    x = common.synth_fast(2, fs, 3 * n)
    for k in range(3):
        blk = x[k * n:(k + 1) * n]
        y_go, y_or = ref.process(go_chain.ch, blk, fs), ochain.process(blk, fs)
        assert np.max(np.abs(y_go - y_or)) <= 1e-12 * max(1.0, np.max(np.abs(y_go))) and np.max(np.abs(y_go)) > 0
        o = G.lit([G.make_slice(0.0, n, None, 'f'), G.make_slice(0.0, n, None, 'f')])
        go_spat.Process(G.lit([G.lit([float(v) for v in y_go], 'f')]), None, o)
        m_go = np.array([o[0].tolist(), o[1].tolist()])
        assert np.max(np.abs(m_go - ospat.process(y_or[None, :]))) <= 1e-12
