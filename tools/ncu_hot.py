#!/usr/bin/env python3
"""Hot SASS regions of an ncu report: tools/ncu_hot.py rep [window] [top]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; W = int(sys.argv[2]) if len(sys.argv) > 2 else 128; TOP = int(sys.argv[3]) if len(sys.argv) > 3 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(raw[1:]))
hdr = rows[0]; rows = rows[1:]
iS = hdr.index("# Samples"); iSrc = hdr.index("Source"); iEx = hdr.index("Instructions Executed")
iExc = hdr.index("L1 Wavefronts Shared Excessive"); iWf = hdr.index("L1 Wavefronts Shared")
tot = sum(int(r[iS] or 0) for r in rows)
print("total samples", tot, "instructions", len(rows))
wins = []
for s in range(0, len(rows), W):
    blk = rows[s:s + W]
    smp = sum(int(r[iS] or 0) for r in blk)
    mix = collections.Counter(r[iSrc].split()[0].split(".")[0] for r in blk if r[iSrc].strip())
    exc = sum(int(float(r[iExc] or 0)) for r in blk); wf = sum(int(float(r[iWf] or 0)) for r in blk)
    ex = sum(int(r[iEx] or 0) for r in blk)
    wins.append((smp, s, mix, exc, wf, ex))
for smp, s, mix, exc, wf, ex in sorted(wins, reverse=True)[:TOP]:
    print("%5.1f%%  instr %6d..%6d  exec %9d  shared wavefronts %9d (excess %9d)  %s" % (100.0 * smp / tot, s, s + W, ex, wf, exc, dict(mix.most_common(6))))
