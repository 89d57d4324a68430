#!/bin/bash
# Build compile-time variants of the streaming kernel HERE (nvcc cross-compiles) so that the GPU box only runs them:
#   tools/build_variants.sh tag1:"-DGCB_STAGES=2" tag2:"-DGCB_CONSUMER_WARPS=16 -DGCB_UNROLL=2" ...
# -> variants/libgcb_<tag>.so, picked up by GCB_LIB=... (gridcount_b200/_cabi.py); tools/sweep_k1.sh runs them.
set -e
cd "$(dirname "$0")/../gridcount_b200/csrc"
make -s > /dev/null
mkdir -p ../../variants
for spec in "$@"; do
  tag=${spec%%:*}; flags=${spec#*:}
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -fmad=false -prec-div=true \
    -prec-sqrt=true -ftz=false -Xcompiler -fPIC,-ffp-contract=off -I../../include $flags -c binning.cu -o /tmp/binning_$tag.o
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/libgcb_$tag.so \
    host_geometry.o grid_io.o /tmp/binning_$tag.o analysis.o xtc_decode.o nccl_reduce.o selftest.o shm_ctl.o local_reduce.o -ldl -lrt
  echo "built variants/libgcb_$tag.so  [$flags]"
done
