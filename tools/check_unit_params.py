#!/usr/bin/env python3
"""Cross-check include/gdg_unit_params.def against the reference's Parameter{} literals.

Build-container only (reads /root/reference/effects/*.go).  Prints every mismatch;
exit status 0 iff the table restates the reference exactly (order, names, physical
units, ranges, defaults, discrete values).
"""
import os, re, sys
REF = os.environ.get("GDG_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = {"SIGNALGENERATOR": "signalgenerator", "NOISEGATE": "noisegate", "BANDPASS": "bandpass",
         "AUTOWAH": "autowah", "AUTOYOY": "autoyoy", "COMPRESSOR": "compressor", "OCTAVER": "octaver",
         "EXCESS": "excess", "FUZZ": "fuzz", "OVERDRIVE": "overdrive", "DISTORTION": "distortion",
         "TONESTACK": "tonestack", "CHORUS": "chorus", "FLANGER": "flanger", "PHASER": "phaser",
         "TREMOLO": "tremolo", "RINGMODULATOR": "ringmodulator", "DELAY": "delay", "REVERB": "reverb",
         "POWERAMP": "poweramp", "CABINET": "cabinet"}

def ref_params(stem):
    src = open(os.path.join(REF, "effects", stem + ".go")).read()
    src = src[src.index("func create"):]
    out = []
    for m in re.finditer(r"Parameter\{(.*?)\n\t\t\t\t\},", src, re.S):
        blk = m.group(1)
        f = lambda k, pat: re.search(k + r":\s*" + pat, blk, re.S)
        name = f("Name", r"\"([^\"]*)\"").group(1)
        typ = f("Type", r"(\w+)").group(1)
        unit = f("PhysicalUnit", r"\"([^\"]*)\"").group(1)
        mn, mx, val, didx = [int(f(k, r"(-?\d+)").group(1)) for k in ("Minimum", "Maximum", "NumericValue", "DiscreteValueIndex")]
        dv = f("DiscreteValues", r"(nil|\[\]string\{.*?\})").group(1)
        if typ == "PARAMETER_TYPE_NUMERIC":
            out.append(("NUM", name, unit, mn, mx, val))
        else:
            out.append(("DISC", name, didx, "|".join(re.findall(r"\"([^\"]*)\"", dv))))
    return out

def my_params():
    src = open(os.path.join(ROOT, "include/gdg_unit_params.def")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    units, params = {}, {}
    for line in src.splitlines():
        line = line.strip()
        m = re.match(r"GDG_UNIT\((\d+), (\w+), \"([^\"]*)\"\)", line)
        if m: units[m.group(2)] = (int(m.group(1)), m.group(3)); params[m.group(2)] = []; continue
        m = re.match(r"GDG_NUM\((\w+), \"([^\"]*)\", \"([^\"]*)\", (-?\d+), (-?\d+), (-?\d+)\)", line)
        if m:
            unit = m.group(3).encode().decode("unicode_escape").encode("latin1").decode("utf8")
            params[m.group(1)].append(("NUM", m.group(2), unit, int(m.group(4)), int(m.group(5)), int(m.group(6)))); continue
        m = re.match(r"GDG_DISC\((\w+), \"([^\"]*)\", (-?\d+), \"([^\"]*)\"\)", line)
        if m: params[m.group(1)].append(("DISC", m.group(2), int(m.group(3)), m.group(4))); continue
        assert not line, line
    return units, params

def main():
    units, mine = my_params()
    eff = open(os.path.join(REF, "effects/effects.go")).read()
    names = re.findall(r"\"([a-z_]+)\",", eff[eff.index("func UnitTypes"):])
    consts = re.findall(r"UNIT_(\w+)", eff[eff.index("UNIT_SIGNALGENERATOR = iota"):eff.index("Mathematical constants")])
    bad = 0
    for code, ident in enumerate(consts):
        if units.get(ident) != (code, names[code]):
            print("unit mismatch", ident, units.get(ident), (code, names[code])); bad += 1
        r = ref_params(FILES[ident])
        if r != mine[ident]:
            print("param mismatch", ident); print("  ref ", r); print("  mine", mine[ident]); bad += 1
    print("checked", len(consts), "units;", bad, "mismatches")
    return 1 if bad else 0

if __name__ == "__main__":
    sys.exit(main())
