#!/bin/bash
# compile-time parameter sweep of the streaming kernel, run on the GPU box: tools/sweep_k1.sh "shape" "cfg1" "cfg2" ...
SHAPES=$1; shift
for cfg in "$@"; do
  rm -f gridcount_b200/csrc/binning.o
  make -s -C gridcount_b200/csrc EXTRA_NVFLAGS="$cfg" > /dev/null 2>&1 || { echo "build failed: $cfg"; continue; }
  for s in $SHAPES; do echo "[$cfg] $(timeout 200 python tools/perf_k1.py $s 2>&1 | grep -E 'stream|routed' | sed 's/grid=.*best/best/' | tr '\n' ' ')"; done
done
rm -f gridcount_b200/csrc/binning.o; make -s -C gridcount_b200/csrc > /dev/null 2>&1
