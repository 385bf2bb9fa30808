#!/bin/bash
# GPU-box visit: the whole -m gpu suite (all failures shown) + scratch timings.  usage: bash tools/visit_tests.sh <tag> [pytest args]
TAG=${1:-v}; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q --durations=15 "$@" > $OUT/${TAG}_gputests.log 2>&1
tail -60 $OUT/${TAG}_gputests.log
timeout 600 python tools/quick_bench.py sedov61 kh1000j kh2000j fb1000j > $OUT/${TAG}_quick.log 2>&1
cat $OUT/${TAG}_quick.log
