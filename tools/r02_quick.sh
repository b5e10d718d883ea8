#!/bin/bash
# quick timings of selected units + CHAIN12 (GDG_ONLY), optional ncu of one unit
mkdir -p gpurun_out
GDG_ONLY=${ONLY:-excess,fuzz,overdrive,distortion} python tools/quick_bench.py 2>&1 | tail -${TAILN:-8}
if [ -n "$PROF" ]; then
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-k_os} -s 2 -c 1 -f -o gpurun_out/${TAG:-r02q}_$PROF python tools/one_unit.py $PROF 1024 > gpurun_out/${TAG:-r02q}_prof_$PROF.log 2>&1
  tail -1 gpurun_out/${TAG:-r02q}_prof_$PROF.log
fi
