#!/usr/bin/env python3
"""SASS op-count summary of libgdg.so per kernel (static instruction counts, not executed counts):
   python tools/sass_summary.py > profiles/r02_sass_opcounts.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "go-dsp-guitar_b200", "libgdg.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
for m in re.finditer(r"Function (\S+):\s*\n\s*(REG:\d+[^\n]*)", res): usage[m.group(1)] = m.group(2).strip()
kern = None; counts = collections.OrderedDict(); arch = set()
for line in out.splitlines():
    m = re.match(r"\s*arch = (\S+)", line)
    if m: arch.add(m.group(1))
    m = re.match(r"\s*Function : (\S+)", line)
    if m: kern = m.group(1); counts[kern] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern: counts[kern][m.group(1)] += 1; counts[kern]["(all)"] += 1
groups = [("FP64", ["DFMA", "DADD", "DMUL", "DSETP", "MUFU"]), ("shared", ["LDS", "STS", "LDSM"]), ("global", ["LDG", "STG", "LDGSTS", "UBLKCP", "SYNCS", "CCTL"]),
          ("local", ["LDL", "STL"]), ("warp/cta", ["SHFL", "BAR", "VOTE"]), ("const", ["LDC", "LDCU", "ULDC"])]
print("# static SASS instruction counts per kernel of %s (arch %s); cuobjdump -sass / -res-usage" % (os.path.basename(lib), ",".join(sorted(arch))))
print("# tensor-core / TMA-tensor mnemonics (UTCHMMA, UTMALDG, HMMA, DMMA): %s" %
      (", ".join("%s=%d" % (k, sum(c[k] for c in counts.values())) for k in ["UTCHMMA", "UTCQMMA", "UTMALDG", "HMMA", "DMMA", "IMMA"])))
for k, c in counts.items():
    print("\n== %s   [%s]   %d instructions" % (k, usage.get(k, "?"), c["(all)"]))
    for g, ops in groups:
        print("  %-9s %s" % (g, "  ".join("%s %d" % (o, c[o]) for o in ops if c[o])))
