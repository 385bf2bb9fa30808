#!/bin/bash
# one single-GPU visit: [tests] + A/B of the given libraries + optional ncu --set full captures
# usage: bash tools/visit.sh <tag> "<pytest args or ->" "<libs for ab_bench>" "<kernel regexes to capture>"
TAG=$1; OUT=gpurun_out; mkdir -p $OUT
if [ "$2" != "-" ]; then timeout 900 python -m pytest $2 -m gpu -q -x 2>&1 | tail -4; fi
[ -n "$3" ] && bash tools/ab_bench.sh $3
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
for K in $4; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o $OUT/${TAG}_prof_$K -f $B > $OUT/${TAG}_ncu_full_$K.log 2>&1
done
