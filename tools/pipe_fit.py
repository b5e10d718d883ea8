#!/usr/bin/env python3
"""prints timings of CHAIN12 for small channel counts, cluster-pipelined oversampled run vs one CTA per channel"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for c in [int(a) for a in sys.argv[1:]] or [1, 16, 64, 96, 128, 148]:
    for env in ({}, {"GDG_NO_OS_PIPE": "1"}):
        e = dict(os.environ, GDG_C=str(c), GDG_ONLY="none", GDG_REPS="40", **env)
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "quick_bench.py")], env=e, capture_output=True, text=True).stdout.strip().splitlines()
        print("C=%4d %-18s %s" % (c, "one CTA/channel" if env else "cluster pipeline", out[-1] if out else "?"), flush=True)
