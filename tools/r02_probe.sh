#!/bin/bash
# round-2 baseline probe on the GPU box: toolchain probe, per-unit timings
mkdir -p gpurun_out
{ echo "== go version"; go version 2>&1; echo "rc=$?"; which go gccgo 2>&1; ls /usr/local/go 2>&1; echo "== nproc"; nproc; lscpu | head -20; echo "== nvidia-smi"; nvidia-smi -L; } > gpurun_out/r02_toolchain_probe.txt 2>&1
python tools/quick_bench.py > gpurun_out/r02a_quick_bench.txt 2>&1
cp gpurun_out/quick_bench.json gpurun_out/r02a_quick_bench.json
tail -30 gpurun_out/r02a_quick_bench.txt
