#!/bin/bash
# ncu --set full captures of k_chain for single-unit chains (1024 channels), reports brought back in gpurun_out/
mkdir -p gpurun_out
for u in ${UNITS:-distortion fuzz overdrive power_amp}; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_chain -s 2 -c 1 -f -o gpurun_out/r02_$u python tools/one_unit.py $u 1024 > gpurun_out/r02_prof_$u.log 2>&1
  tail -2 gpurun_out/r02_prof_$u.log
done
ls -la gpurun_out/*.ncu-rep
