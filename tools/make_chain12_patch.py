#!/usr/bin/env python3
"""Writes examples/chain12.patch.json: BASELINE config 4 (8 channels of CHAIN12 with the reference's "Guitar: American Vintage
(Center)" response, spatializer -70 .. +70 degrees in 20-degree steps, distance 1 + 0.25 c m, level 1) as a
persistence.Configuration - the file format the reference's own "save patch" button writes (controller.go:1195-1370).
The chains are built through the Chain API of the test-only emulation backend (no GPU needed; the parameter store is the
host runtime's, the same code the product library runs) and written with go-dsp-guitar_b200/patch.py.
    python tools/make_chain12_patch.py [--check]"""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import common  # noqa: E402
from common import gdg  # noqa: E402

patch = importlib.import_module("go-dsp-guitar_b200.patch")
OUT = os.path.join(ROOT, "examples", "chain12.patch.json")
IR = "Guitar: American Vintage (Center)"
CHANNELS = 8


def build(lib):
    _ob, pb = common.make_banks(lib)
    chains = [common.build_chain(gdg.Chain(pb, lib=lib), common.CHAIN12, IR) for _ in range(CHANNELS)]
    spat = gdg.Spatializer(CHANNELS, lib=lib)
    for c in range(CHANNELS):
        spat.set_azimuth(c, -70.0 + 20.0 * c)
        spat.set_distance(c, 1.0 + 0.25 * c)
        spat.set_level(c, 1.0)
    return chains, spat


def text(lib):
    chains, spat = build(lib)
    return json.dumps(patch.persist(chains, spat, frames_per_period=8192), indent=1) + "\n"


if __name__ == "__main__":
    t = text(common.backend_lib("emu"))
    if "--check" in sys.argv:
        sys.exit(0 if open(OUT).read() == t else 1)
    open(OUT, "w").write(t)
    print(OUT, len(t), "bytes")
