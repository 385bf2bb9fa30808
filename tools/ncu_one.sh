#!/bin/bash
# usage: bash tools/ncu_one.sh <tag> <workload> <kernel regex> [skip]  -> gpurun_out/<tag>_<kernel>_<workload>.ncu-rep (+ summary txt)
TAG=$1; WL=$2; K=$3; SKIP=${4:-2}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -o $OUT/${TAG}_${K}_${WL} -f python tools/quick_bench.py $WL > $OUT/${TAG}_ncu_${K}_${WL}.log 2>&1
python tools/ncu_summary.py $OUT/${TAG}_${K}_${WL}.ncu-rep > $OUT/${TAG}_${K}_${WL}_ncu_full.txt 2>&1
head -60 $OUT/${TAG}_${K}_${WL}_ncu_full.txt
