#!/bin/bash
# A/B of library builds on the scratch timings.  usage: bash tools/visit_ab.sh <tag> "<workloads>" lib1.so lib2.so ...
TAG=$1; WL=$2; shift; shift
OUT=gpurun_out; mkdir -p $OUT
for so in "$@"; do
  echo "=== $so"
  MLH_GPU_LIB=$PWD/meshlesshydro_b200/$so timeout 400 python tools/quick_bench.py $WL 2>&1 | grep -E "N=|k[0-9]"
done 2>&1 | tee $OUT/${TAG}_ab.log
