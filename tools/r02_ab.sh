#!/bin/bash
# A/B: current libgdg.so vs tools/ab/libgdg_prev.so on the same box
cp go-dsp-guitar_b200/libgdg.so /tmp/libgdg_cur.so
for v in cur prev cur prev; do
  if [ $v = prev ]; then cp tools/ab/libgdg_prev.so go-dsp-guitar_b200/libgdg.so; else cp /tmp/libgdg_cur.so go-dsp-guitar_b200/libgdg.so; fi
  echo "== $v"; GDG_ONLY=${ONLY:-excess,fuzz,overdrive,distortion} python tools/quick_bench.py 2>&1 | tail -${TAILN:-5}
done
cp /tmp/libgdg_cur.so go-dsp-guitar_b200/libgdg.so
