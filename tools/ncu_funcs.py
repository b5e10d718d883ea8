#!/usr/bin/env python3
"""Per-device-function share of the warp-stall samples of k_chain in an ncu report.
usage: tools/ncu_funcs.py rep.ncu-rep [libgdg.so]   (the .so must be the build that was profiled)"""
import csv, os, re, subprocess, sys, tempfile, collections
rep = sys.argv[1]
so = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "go-dsp-guitar_b200", "libgdg.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin") and "host" not in f][0]
sym = subprocess.run(["readelf", "-sW", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
funcs = []
for line in sym.splitlines():
    m = re.match(r"\s*\d+:\s+([0-9a-f]+)\s+(\d+)\s+FUNC\s+\S+\s+\S+\s+(\d+)\s+(\S+)", line)
    if m and "k_chain" in m.group(4):
        name = m.group(4).split("$")[-1] if "$" in m.group(4) else "k_chain(main)"
        funcs.append((int(m.group(1), 16), int(m.group(2)), name))
funcs.sort()
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(raw[1:])); hdr = rows[0]; rows = rows[1:]
iS = hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.Counter(); stalls = collections.defaultdict(collections.Counter)
def func_of(off):
    name = "k_chain(main)"
    for a, sz, nm in funcs:
        if a <= off < a + sz: name = nm
    return name
for i, r in enumerate(rows):
    n = int(r[iS] or 0)
    if not n: continue
    f = func_of(16 * i)
    agg[f] += n
    for c in stall_cols: stalls[f][hdr[c][6:]] += int(float(r[c] or 0))
tot = sum(agg.values())
import subprocess as sp
def dem(n):
    try: return sp.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
    except Exception: return n
for f, n in agg.most_common(25):
    top = ", ".join("%s %d%%" % (k, 100 * v / max(1, sum(stalls[f].values()))) for k, v in stalls[f].most_common(4))
    print("%5.1f%%  %-46s %s" % (100.0 * n / tot, dem(f)[:46], top))

# optional: hot 128-instruction windows inside one function (argv[3] = substring of its name)
if len(sys.argv) > 3:
    want = sys.argv[3]
    for a, sz, nm in funcs:
        if want in dem(nm):
            i0, i1 = a // 16, (a + sz) // 16
            ftot = sum(int(rows[i][iS] or 0) for i in range(i0, min(i1, len(rows))))
            print("---- %s: %d instructions, %d samples" % (dem(nm), i1 - i0, ftot))
            iSrc = hdr.index("Source")
            for s0 in range(i0, i1, 128):
                blk = rows[s0:min(s0 + 128, i1)]
                smp = sum(int(r[iS] or 0) for r in blk)
                mix = collections.Counter(r[iSrc].split()[0].split(".")[0] for r in blk if r[iSrc].strip())
                st = collections.Counter()
                for r in blk:
                    for c in stall_cols: st[hdr[c][6:]] += int(float(r[c] or 0))
                print("%5.1f%% +%5d %s | %s" % (100.0 * smp / max(1, ftot), s0 - i0, dict(mix.most_common(5)), ", ".join("%s %d%%" % (k, 100 * v / max(1, sum(st.values()))) for k, v in st.most_common(3))))
