#!/usr/bin/env python3
"""Warp-stall samples of a kernel per CUDA source line (needs -lineinfo and --import-source on).
usage: tools/ncu_lines.py rep.ncu-rep [top=40] [file-substring]"""
import collections, csv, os, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40; only = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout.splitlines()
agg = collections.Counter(); stalls = collections.defaultdict(collections.Counter); text = {}; execs = collections.Counter(); ninst = collections.Counter()
conf = collections.Counter()
cur = None; hdr = None; last = -1
for row in csv.reader(raw):
    if not row: continue
    if row[0] == "File Path": cur = os.path.basename(row[1]); continue
    if row[0] == "Function Name": continue
    if row[0] == "Line No": hdr = row; iS = hdr.index("# Samples"); iE = hdr.index("Instructions Executed"); iX = hdr.index("L1 Wavefronts Shared Excessive"); sc = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]; continue
    if hdr is None or len(row) < len(hdr): continue
    try: n = int(row[iS] or 0)
    except ValueError: continue
    if row[0].isdigit(): last = int(row[0])
    key = (cur, last)
    if row[1].strip(): text[key] = row[1].strip()
    agg[key] += n; execs[key] += int(float(row[iE] or 0)); ninst[key] += 1
    conf[key] += int(float(row[iX] or 0))
    for c in sc: stalls[key][hdr[c][6:]] += int(float(row[c] or 0))
tot = sum(agg.values())
print("total samples %d" % tot)
byfile = collections.Counter()
for (f, l), n in agg.items(): byfile[f] += n
print("by file:", ", ".join("%s %.1f%%" % (f, 100.0 * n / tot) for f, n in byfile.most_common(6)))
for key, n in [kv for kv in agg.most_common(400) if only in kv[0][0]][:top]:
    st = stalls[key]; s = sum(st.values()) or 1
    print("%5.1f%% %-18s:%-4d x%-9d i%-4d bc%-9d %-34s | %s" % (100.0 * n / tot, key[0][:18], key[1], execs[key], ninst[key], conf[key],
          ", ".join("%s %d%%" % (k, 100 * v / s) for k, v in st.most_common(3)), text.get(key, "")[:90]))
