#!/bin/bash
# Run on the GPU box (through gpurun): ncu captures behind profiles/ for every kernel on a default path.
#   tools/profile_all.sh TAG            -> gpurun_out/{launches,raw,src}_<name>_TAG.csv
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
make -s -C host libgcbhost.so
full() {   # name kernel-regex skip command...
  local name=$1 regex=$2 skip=$3; shift 3
  # (the .ncu-rep stays on the box: gpurun brings back at most 64 MiB; the two CSV pages are what gets summarised)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -f -o /tmp/prof_${name}_$TAG "$@" > $OUT/ncu_full_${name}_$TAG.log 2>&1
  echo "ncu full $name rc=$?"
  ncu -i /tmp/prof_${name}_$TAG.ncu-rep --page raw --csv > $OUT/raw_${name}_$TAG.csv 2>/dev/null
  ncu -i /tmp/prof_${name}_$TAG.ncu-rep --page source --csv > $OUT/src_${name}_$TAG.csv 2>/dev/null
  rm -f /tmp/prof_${name}_$TAG.ncu-rep
}
# the launch list of the bench command itself (shares of the step)
python bench.py --steps 2 --warmup 3 > $OUT/bench_plain_$TAG.log 2> $OUT/bench_plain_$TAG.err || { echo "bench failed"; tail -5 $OUT/bench_plain_$TAG.err; }
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_bench_$TAG.csv python bench.py --steps 2 --warmup 3 > $OUT/ncu_launches_bench_$TAG.log 2>&1
echo "ncu launches bench rc=$?"
export PERF_KERNELS=auto
full c2 bin_stream 4 python tools/perf_k1.py c2
full c3 bin_stream 4 python tools/perf_k1.py c3
full c4 bin_stream 4 python tools/perf_k1.py c4
full c4apply packed_apply_mid 2 python tools/perf_k1.py c4
full p1 bin_stream 4 python tools/perf_k1.py pore
full c1 bin_stream 4 python tools/perf_k1.py c1long
full c3apply packed_apply 2 python tools/perf_k1.py c3
unset PERF_KERNELS
full xtcskim xtc_skim 3 python tools/perf_xtc.py water 60000 64 16
full xtcbin xtc_bin 3 python tools/perf_xtc.py water 60000 64 16
full k3chain an_chain_grid 1 python tools/perf_k3.py c3
full k3slice an_slice_scan 1 python tools/perf_k3.py c3
full k3rdf an_rdf_scan 1 python tools/perf_k3.py c3
ls -la $OUT/*_$TAG.csv | wc -l
