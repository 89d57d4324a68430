#!/bin/bash
# run on the GPU box: add_xtc throughput against batch size and segment decoder (fused / two-step)
for kind in uniform water; do
  for b in default 64 128 256 512 1024; do
    for mode in fused unfused; do
      if [ $b = default ]; then unset GCB_XTC_BATCH; else export GCB_XTC_BATCH=$b; fi
      if [ $mode = unfused ]; then export GCB_XTC_UNFUSED=1; else unset GCB_XTC_UNFUSED; fi
      echo "[$kind batch=$b $mode] $(GCB_TRACE=1 python tools/perf_xtc.py $kind 60000 64 64 2>&1 | grep -E 'add_xtc' | tail -2 | tr '\n' ' ')"
    done
  done
done
