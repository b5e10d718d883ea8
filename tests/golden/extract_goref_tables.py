#!/usr/bin/env python3
"""Extract the anti-aliasing FIR literals of /root/reference/oversampling/oversampling.go (CreateOversamplerDecimator,
x2: 77 taps, x4: 155 taps) into tests/golden/goref_tables.json for tests/goref.py.  Run in the build container only
(the reference tree does not travel to the GPU box); the JSON it writes is committed."""
import json
import os
import re

SRC = "/root/reference/oversampling/oversampling.go"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    text = open(SRC).read()
    body = text[text.index("func CreateOversamplerDecimator"):]
    lists = re.findall(r"coeffs := \[\]float64\{(.*?)\n\t\t\}", body, re.S)
    assert len(lists) == 2, len(lists)
    tabs = [[float(v) for v in re.findall(r"-?\d+\.\d+(?:e-?\d+)?", lst)] for lst in lists]
    assert len(tabs[0]) == 77 and len(tabs[1]) == 155, (len(tabs[0]), len(tabs[1]))
    with open(os.path.join(HERE, "goref_tables.json"), "w") as f:
        json.dump({"source": "oversampling/oversampling.go:239-317,357-513", "aa2": tabs[0], "aa4": tabs[1]}, f)
    print("wrote goref_tables.json: %d + %d taps" % (len(tabs[0]), len(tabs[1])))


if __name__ == "__main__":
    main()
