#!/bin/bash
# Run on the GPU box (through gpurun): the two ncu passes of /opt/skills/guides/B200_PROFILING.md on tools/perf_k1.py
# for one shape/kernel.  Usage: tools/gpu_profile.sh TAG SHAPE KERNEL KERNEL_REGEX [launch-skip]
TAG=$1; SHAPE=$2; KERN=$3; REGEX=$4; SKIP=${5:-4}
OUT=gpurun_out
mkdir -p $OUT
PERF_KERNELS=$KERN timeout 300 python tools/perf_k1.py $SHAPE > $OUT/perf_$TAG.log 2>&1 || { echo "perf failed"; tail -5 $OUT/perf_$TAG.log; exit 1; }
cat $OUT/perf_$TAG.log
PERF_KERNELS=$KERN timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
  --log-file $OUT/launches_$TAG.csv python tools/perf_k1.py $SHAPE > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
PERF_KERNELS=$KERN timeout 900 ncu --set full --clock-control none --import-source on -k regex:$REGEX -s $SKIP -c 1 \
  -f -o $OUT/prof_$TAG python tools/perf_k1.py $SHAPE > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
ncu -i $OUT/prof_$TAG.ncu-rep --page raw --csv > $OUT/raw_$TAG.csv 2>/dev/null
ncu -i $OUT/prof_$TAG.ncu-rep --page source --csv > $OUT/src_$TAG.csv 2>/dev/null
