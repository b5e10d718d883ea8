#!/bin/bash
# round-2 evidence: launch list of the bench command (shares), ncu --set full of the three CHAIN12 segment kernels
mkdir -p gpurun_out
GDG_BENCH_BLOCKS=4 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG:-r02k}_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${TAG:-r02k}_bench_under_ncu.log 2>&1
tail -c 300 gpurun_out/${TAG:-r02k}_bench_under_ncu.log
GDG_WAYS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_chain|k_os' -s 6 -c 3 -f -o gpurun_out/${TAG:-r02k}_chain12 python tools/one_unit.py CHAIN12 1024 > gpurun_out/${TAG:-r02k}_prof_chain12.log 2>&1
tail -2 gpurun_out/${TAG:-r02k}_prof_chain12.log
ls -la gpurun_out/${TAG:-r02k}*
