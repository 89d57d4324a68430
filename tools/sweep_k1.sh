#!/bin/bash
# run on the GPU box: tools/sweep_k1.sh "shapes" "kernels" -> one line per (variant, shape) for every variants/libgcb_*.so
SHAPES=$1; KERN=${2:-stream}
for lib in gridcount_b200/libgridcount_b200.so variants/libgcb_*.so; do
  [ -f "$lib" ] || continue
  for s in $SHAPES; do
    echo "[$(basename $lib)] $(GCB_LIB=$PWD/$lib PERF_KERNELS=$KERN timeout 200 python tools/perf_k1.py $s 2>&1 | grep -E 'best' | sed 's/grid=.*best/best/' | tr '\n' ' ')"
  done
done
