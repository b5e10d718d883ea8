#!/bin/bash
# Is the device code of go-dsp-guitar_b200/libgdg.so the device code of COMMIT?  Rebuilds COMMIT's csrc/ + include/ in a scratch
# directory with the product's own nvcc line and compares `cuobjdump -sass` of the two libraries (address comments and the
# source-path identifier lines stripped).    tools/sass_same_as.sh 775f6e2
set -e
COMMIT=${1:?commit}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
git -C "$ROOT" archive "$COMMIT" go-dsp-guitar_b200/csrc include | tar -C "$TMP" -xf -
(cd "$TMP" && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O2,-Wall -shared \
    -o old.so go-dsp-guitar_b200/csrc/gdg_kernels.cu go-dsp-guitar_b200/csrc/gdg_host.cpp -lcudart)
norm() { cuobjdump -sass "$1" | grep -v "^Fatbin\|^\s*$\|^identifier" | sed 's#/\*[0-9a-f]\{4,\}\*/##'; }
A=$(norm "$TMP/old.so" | md5sum | cut -d' ' -f1); B=$(norm "$ROOT/go-dsp-guitar_b200/libgdg.so" | md5sum | cut -d' ' -f1)
LINES=$(norm "$ROOT/go-dsp-guitar_b200/libgdg.so" | wc -l)
rm -rf "$TMP"
echo "$COMMIT: $A"; echo "worktree: $B ($LINES lines of SASS)"
[ "$A" = "$B" ] && echo "device code identical" || { echo "device code DIFFERS"; exit 1; }
