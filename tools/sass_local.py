#!/usr/bin/env python3
"""Local-memory (LDL/STL) instructions of k_chain per device function, from the SASS of libgdg.so.
usage: tools/sass_local.py [libgdg.so]"""
import collections, os, re, subprocess, sys, tempfile
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "go-dsp-guitar_b200", "libgdg.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
cubin = os.path.join(tmp, [f for f in os.listdir(tmp) if f.endswith(".cubin") and "host" not in f][0])
sym = subprocess.run(["readelf", "-sW", cubin], capture_output=True, text=True).stdout
funcs = []
for line in sym.splitlines():
    m = re.match(r"\s*\d+:\s+([0-9a-f]+)\s+(\d+)\s+FUNC\s+\S+\s+\S+\s+(\d+)\s+(\S+)", line)
    if m and "k_chain" in m.group(4):
        funcs.append((int(m.group(1), 16), int(m.group(2)), m.group(4).split("$")[-1] if "$" in m.group(4) else "k_chain(main)"))
full = subprocess.run(["cuobjdump", "-sass", os.path.abspath(so)], capture_output=True, text=True).stdout
sass, on = [], False
for line in full.splitlines():
    if "Function :" in line: on = "k_chain" in line
    elif on: sass.append(line)
sass = "\n".join(sass)
cnt = collections.Counter(); tot = collections.Counter()
for line in sass.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if not m: continue
    off = int(m.group(1), 16)
    name = "k_chain(main)"
    for a, sz, nm in funcs:
        if a <= off < a + sz: name = nm
    tot[name] += 1
    if re.search(r"\b(LDL|STL)\b", m.group(2)): cnt[name] += 1
dem = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0]
for n, c in cnt.most_common():
    print("%5d local of %6d  %s" % (c, tot[n], dem(n)))
