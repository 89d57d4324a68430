#!/bin/bash
# Run on the GPU box (through gpurun): tests, plain bench, then the two ncu passes of
# /opt/skills/guides/B200_PROFILING.md on the same bench command.  Usage: tools/gpu_profile.sh TAG [skip_tests]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
if [ -z "$2" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1
  echo "pytest rc=$?" | tee -a $OUT/pytest_$TAG.log
  tail -3 $OUT/pytest_$TAG.log
fi
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench_$TAG.log 2>&1 || { echo "bench failed"; tail -20 $OUT/bench_$TAG.log; exit 1; }
tail -1 $OUT/bench_$TAG.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $OUT/launches_$TAG.csv python bench.py --steps 2 --warmup 3 > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bin_stream_kernel -s 3 -c 1 \
  -f -o $OUT/prof_k1_$TAG python bench.py --steps 2 --warmup 3 > $OUT/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
ncu -i $OUT/prof_k1_$TAG.ncu-rep --page raw --csv > $OUT/raw_$TAG.csv 2>/dev/null
ncu -i $OUT/prof_k1_$TAG.ncu-rep --page source --csv > $OUT/src_$TAG.csv 2>/dev/null
for s in c1 c2 c2all c3 c4; do timeout 300 python tools/perf_k1.py $s; done 2>&1 | tee $OUT/perf_$TAG.log
