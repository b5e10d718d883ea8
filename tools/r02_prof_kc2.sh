#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_chain' -s 5 -c 1 -f -o gpurun_out/${TAG:-r02l}_kchain2 python tools/one_unit.py CHAIN12 1024 > gpurun_out/${TAG:-r02l}_prof.log 2>&1
tail -1 gpurun_out/${TAG:-r02l}_prof.log
