#!/bin/bash
mkdir -p gpurun_out
for u in ${UNITS:-distortion fuzz}; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-k_os} -s 2 -c 1 -f -o gpurun_out/r02d_$u python tools/one_unit.py $u 1024 > gpurun_out/r02d_prof_$u.log 2>&1
  tail -2 gpurun_out/r02d_prof_$u.log
done
