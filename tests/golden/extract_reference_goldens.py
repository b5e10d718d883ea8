#!/usr/bin/env python3
"""Extract the reference's own golden vectors into committed JSON / NPZ fixtures.

Run in the BUILD container only (it reads /root/reference, which does not exist
on the GPU box):

    python tests/golden/extract_reference_goldens.py

What it reads (and nothing else):
  * slice literals (``name := [][]float64{...}`` etc.) inside the ``Test*``
    functions of the reference's ``*_test.go`` files -> ``reference_goldens.json``
      fft/fft_test.go, oversampling/oversampling_test.go, resample/resample_test.go,
      random/random_test.go, circular/circular_test.go
  * wave/wave_test.go: per sample format the export vector (samples -> expected WAV bytes; only the data chunk
    payload is kept) and the import vector (WAV bytes -> expected samples, with the test's tolerance)
    -> ``wave_goldens.json``
  * the 24 impulse-response WAVs named by ``ir/index.json`` (16-bit PCM mono,
    96 kHz) -> ``ir_bank.npz`` holding the raw int16 frames + names + dB
    compensation (data fixtures for power-amp parity, SURVEY.md §2 "ir/")
  * the last 96000 frames (= tuner ring size, tuner/tuner.go:16) of the six
    tuner test recordings tuner/samples/*.wav -> ``tuner_tail_<note>.npz``

No Go source is copied: only numeric test data is re-serialised.
"""
import json
import os
import re
import struct
import sys

import numpy as np

REF = os.environ.get("GDG_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def strip_comments(src):
    return re.sub(r"/\*.*?\*/", "", src, flags=re.S)


def split_functions(src):
    """Return {func_name: body} for top-level ``func Test...``."""
    out = {}
    for m in re.finditer(r"^func (Test\w+)\(t \*testing\.T\) \{$", src, flags=re.M):
        start = m.end()
        depth = 1
        i = start
        while depth and i < len(src):
            c = src[i]
            if c == "{":
                depth += 1
            elif c == "}":
                depth -= 1
            i += 1
        out[m.group(1)] = src[start:i - 1]
    return out


NUM = r"[-+]?(?:0x[0-9a-fA-F]+|\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+)"


def parse_scalar(tok, kind):
    tok = tok.strip()
    if kind in ("uint64", "uint32", "int"):
        return int(tok, 0)
    return float(tok)


def parse_literal(text, typ):
    """text is the inside of the outermost braces of a literal of Go type typ."""
    if typ.startswith("[][]"):
        inner = typ[2:]
        rows = []
        for m in re.finditer(re.escape(inner) + r"\{(.*?)\}", text, flags=re.S):
            rows.append(parse_literal(m.group(1), inner))
        return rows
    kind = typ[2:]
    if kind == "complex128":
        return [[float(a), float(b)] for a, b in
                re.findall(r"complex\(\s*(%s)\s*,\s*(%s)\s*\)" % (NUM, NUM), text)]
    return [parse_scalar(t, kind) for t in re.findall(NUM, text)]


def extract_literals(body):
    out = {}
    pat = re.compile(r"(\w+) := ((?:\[\])+(?:float64|complex128|uint64|uint32|int))\{")
    for m in pat.finditer(body):
        name, typ = m.group(1), m.group(2)
        i = m.end()
        depth = 1
        while depth:
            c = body[i]
            if c == "{":
                depth += 1
            elif c == "}":
                depth -= 1
            i += 1
        val = parse_literal(body[m.end():i - 1], typ)
        key = name
        n = 2
        while key in out:          # same name reused later in one function
            key = "%s_%d" % (name, n)
            n += 1
        out[key] = val
    return out


def wav_chunks(buf):
    assert buf[:4] == b"RIFF" and buf[8:12] == b"WAVE"
    pos = 12
    chunks = {}
    while pos + 8 <= len(buf):
        cid = buf[pos:pos + 4]
        size = struct.unpack("<I", buf[pos + 4:pos + 8])[0]
        chunks[cid] = buf[pos + 8:pos + 8 + size]
        pos += 8 + size + (size & 1)
    return chunks


def main():
    goldens = {}
    for rel in ["fft/fft_test.go", "oversampling/oversampling_test.go",
                "resample/resample_test.go", "random/random_test.go",
                "circular/circular_test.go"]:
        src = strip_comments(open(os.path.join(REF, rel)).read())
        funcs = split_functions(src)
        goldens[rel] = {fn: extract_literals(body) for fn, body in funcs.items()}
        print(rel, {fn: list(v) for fn, v in goldens[rel].items()})
    with open(os.path.join(HERE, "reference_goldens.json"), "w") as f:
        json.dump(goldens, f, indent=0)

    # --- wave codec vectors (wave/wave_test.go) ---
    wsrc = strip_comments(open(os.path.join(REF, "wave/wave_test.go")).read())
    wf = split_functions(wsrc)
    wave = {}
    for fmt in ["PCM8", "PCM16", "PCM24", "PCM32", "IEEE32", "IEEE64"]:
        exp, imp = wf["TestExport%sMono" % fmt], wf["TestImport%sMono" % fmt]
        tobytes = lambda body, name: [int(x, 16) for x in re.findall(r"0x[0-9a-fA-F]{2}", re.search(name + r" := \[\]byte\{(.*?)\}", body, re.S).group(1))]
        samples = extract_literals(exp)["samples"]
        out = tobytes(exp, "expectedOutput")
        buf = tobytes(imp, "buf")
        expected = extract_literals(imp)["expectedSamples"]
        tol = float(re.search(r"areSlicesClose\(samples, expectedSamples, ([0-9.e+-]+)\)", imp).group(1))

        def payload(b):
            pos = 12
            while pos + 8 <= len(b):
                cid, size = bytes(b[pos:pos + 4]), struct.unpack("<I", bytes(b[pos + 4:pos + 8]))[0]
                if cid == b"data":
                    return b[pos + 8:pos + 8 + size]
                pos += 8 + size + (size & 1)
            raise ValueError("no data chunk")
        wave[fmt] = {"export_samples": samples, "export_data": payload(out), "import_data": payload(buf),
                     "import_samples": expected, "import_tolerance": tol}
        print("wave", fmt, len(samples), len(wave[fmt]["export_data"]), tol)
    with open(os.path.join(HERE, "wave_goldens.json"), "w") as f:
        json.dump(wave, f)

    # --- impulse-response bank (data fixtures) ---
    index = json.load(open(os.path.join(REF, "ir/index.json")))
    names, comps, frames = [], [], {}
    for k, d in enumerate(index):
        buf = open(os.path.join(REF, d["Path"]), "rb").read()
        ch = wav_chunks(buf)
        fmt = struct.unpack("<HHIIHH", ch[b"fmt "][:16])
        assert fmt[0] == 1 and fmt[1] == 1 and fmt[2] == 96000 and fmt[5] == 16, (d, fmt)
        pcm = np.frombuffer(ch[b"data"], dtype="<i2").copy()
        names.append(d["Name"])
        comps.append(int(d["Compensation"]))
        frames["pcm_%02d" % k] = pcm
    np.savez_compressed(os.path.join(HERE, "ir_bank.npz"),
                        names=np.array(names), compensation=np.array(comps, dtype=np.int32),
                        sample_rate=np.array(96000, dtype=np.uint32), **frames)
    print("ir_bank:", len(names), "responses,", sum(len(v) for v in frames.values()), "frames")

    # --- tuner recordings: only the last 96000 frames matter (circular.Enqueue
    #     of >= ring size keeps the tail, circular/circular.go:38-44) ---
    for note in ["D2", "A2", "D3", "G3", "H3", "E4"]:
        buf = open(os.path.join(REF, "tuner/samples/%s.wav" % note), "rb").read()
        ch = wav_chunks(buf)
        fmt = struct.unpack("<HHIIHH", ch[b"fmt "][:16])
        assert fmt[0] == 3 and fmt[1] == 1 and fmt[5] == 64, fmt
        x = np.frombuffer(ch[b"data"], dtype="<f8")
        np.savez_compressed(os.path.join(HERE, "tuner_tail_%s.npz" % note),
                            samples=x[-96000:].copy(),
                            sample_rate=np.array(fmt[2], dtype=np.uint32))
        print("tuner", note, fmt[2], len(x))

    # --- level meter: level/level_test.go TestMeters feeds one second of a 1 Hz sine (amplitude 1 and 0.5) at
    #     96 kHz in ONE Process call and expects integer dB read-outs before and after SetEnabled(false) ---
    lsrc = strip_comments(open(os.path.join(REF, "level/level_test.go")).read())
    rate = int(re.search(r"DEFAULT_SAMPLE_RATE\s*=\s*(\d+)", lsrc).group(1))
    lv = [int(v) for v in re.findall(r"expectedLevel := int32\((-?\d+)\)", lsrc)]
    pk = [int(v) for v in re.findall(r"expectedPeak := int32\((-?\d+)\)", lsrc)]
    assert len(lv) == 4 and len(pk) == 4 and "0.5 * elemA" in lsrc and "TWO_PI * t" in lsrc, (lv, pk)
    level = {"source": "level/level_test.go", "sample_rate": rate, "samples": rate, "signal": "sin(2 pi i / sample_rate)",
             "amplitudes": [1.0, 0.5],
             "enabled": {"level": lv[0:2], "peak": pk[0:2]},
             "after_disable": {"level": lv[2:4], "peak": pk[2:4]}}
    with open(os.path.join(HERE, "level_goldens.json"), "w") as f:
        json.dump(level, f, indent=1)
    print("level:", level["enabled"], level["after_disable"])


if __name__ == "__main__":
    sys.exit(main())
