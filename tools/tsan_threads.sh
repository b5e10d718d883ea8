#!/bin/bash
# The host runtime's threading contract under ThreadSanitizer: the test-only emulation library (gdg_host.cpp + the stage
# source as host code) is built with -fsanitize=thread in a scratch copy of the tree and the tests that drive the library from
# two threads (setters vs processing, spatializer setters, tuner polls, one processing thread per channel) run there.  Prints TSan's report count.
set -e
SRC=$(cd "$(dirname "$0")/.." && pwd)
DST=${TSAN_DIR:-/tmp/gdg_tsan}
rm -rf "$DST"; mkdir -p "$DST"
tar -C "$SRC" --exclude=.git --exclude=gpurun_out --exclude='*.so' --exclude=__pycache__ --exclude=profiles -cf - . | tar -C "$DST" -xf -
g++ -O1 -g -std=c++17 -fPIC -shared -Wall -Wno-unused-function -Wno-unknown-pragmas -ffp-contract=off -fsanitize=thread \
    -o "$DST/tests/emu/libgdg_emu.so" "$DST/tests/emu/emu_backend.cpp" "$DST/go-dsp-guitar_b200/csrc/gdg_host.cpp" -lm -lpthread
make -C "$DST/oracle" -s
cp "$SRC/go-dsp-guitar_b200/libgdg.so" "$DST/go-dsp-guitar_b200/" 2>/dev/null || true
touch "$DST/tests/emu/libgdg_emu.so" "$DST/oracle/liborc.so" "$DST/go-dsp-guitar_b200/libgdg.so"
cd "$DST"
LD_PRELOAD=$(gcc -print-file-name=libtsan.so) TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0" \
    python -m pytest tests/test_next_rows.py tests/test_edge_cases.py -q -m "not gpu" -p no:cacheprovider -k "second_thread or race_free or one_thread_per_channel" "$@" 2>&1 | tee /tmp/gdg_tsan.log | tail -3
echo "ThreadSanitizer reports: $(grep -o 'reported [0-9]* warnings' /tmp/gdg_tsan.log || echo 'reported 0 warnings')"
