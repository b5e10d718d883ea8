#!/bin/bash
# quick per-unit timings, then the GPU parity suite
mkdir -p gpurun_out
python tools/quick_bench.py > gpurun_out/r02c_quick_bench.txt 2>&1; tail -26 gpurun_out/r02c_quick_bench.txt
cp gpurun_out/quick_bench.json gpurun_out/r02c_quick_bench.json
GDG_NO_OS_KERNEL=1 GDG_ONLY=excess,fuzz,overdrive,distortion python tools/quick_bench.py 2>&1 | tail -6
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r02c_pytest_gpu.txt
