#!/bin/bash
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
for so in libmlh_gpu.so "$@"; do
  echo "=== $so"
  MLH_GPU_LIB=$PWD/meshlesshydro_b200/$so timeout 300 python tools/quick_bench.py sedov61 kh1000j 2>&1 | grep -E "N=|k4b1|k4a|k4b3"
done
