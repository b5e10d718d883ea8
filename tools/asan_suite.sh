#!/bin/bash
# Host runtime + oracle under AddressSanitizer / UBSan: copies the tree to a scratch directory, builds the test-only
# emulation library (gdg_host.cpp + the stage source as host code) and the oracle with -fsanitize=address,undefined and
# runs the CPU suite there.  The product library is not involved (nvcc device code cannot be sanitised this way;
# compute-sanitizer covers it on the GPU box, tools/sanitize.py).      tools/asan_suite.sh [pytest args]
set -e
SRC=$(cd "$(dirname "$0")/.." && pwd)
DST=${ASAN_DIR:-/tmp/gdg_asan}
rm -rf "$DST"; mkdir -p "$DST"
tar -C "$SRC" --exclude=.git --exclude=gpurun_out --exclude='*.so' --exclude=__pycache__ --exclude=profiles -cf - . | tar -C "$DST" -xf -
SAN="-fsanitize=address,undefined -fno-omit-frame-pointer -g"
g++ -O1 -std=c++17 -fPIC -shared -Wall -Wno-unused-function -Wno-unknown-pragmas -ffp-contract=off $SAN \
    -o "$DST/tests/emu/libgdg_emu.so" "$DST/tests/emu/emu_backend.cpp" "$DST/go-dsp-guitar_b200/csrc/gdg_host.cpp" -lm -lpthread
gcc -O1 -std=gnu11 -fPIC -Wall -Wextra -ffp-contract=off -fno-fast-math $SAN -shared \
    -o "$DST/oracle/liborc.so" "$DST"/oracle/orc_dsp.c "$DST"/oracle/orc_units.c "$DST"/oracle/orc_mix.c "$DST"/oracle/orc_bench.c -lm -lpthread
cp "$SRC/go-dsp-guitar_b200/libgdg.so" "$DST/go-dsp-guitar_b200/" 2>/dev/null || true      # the symbol-table test only dlopens it (nvcc does not run under the sanitizer runtime)
touch "$DST/tests/emu/libgdg_emu.so" "$DST/oracle/liborc.so" "$DST/go-dsp-guitar_b200/libgdg.so"          # newer than every source: the suite must not rebuild them
cd "$DST"
export LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so)
export ASAN_OPTIONS=detect_leaks=0:abort_on_error=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1
exec python -m pytest tests -q -m "not gpu" -x -p no:cacheprovider "$@"
