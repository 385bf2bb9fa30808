#!/bin/bash
TAG=${1:-v}; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 120 > $OUT/${TAG}_parity.log 2>&1
tail -15 $OUT/${TAG}_parity.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest "tests/test_gpu_parity.py::test_one_step_matches_oracle" -m gpu -q -x -k "kh_random_50-1 or sedov_21-0 or fb_jitter_60-0" > $OUT/${TAG}_sanitizer.log 2>&1
echo "sanitizer rc=$?"; grep -E "ERROR SUMMARY|Invalid|passed|failed" $OUT/${TAG}_sanitizer.log | head -20
timeout 1800 python -m pytest tests -m gpu -q --timeout 300 --durations=8 > $OUT/${TAG}_gputests.log 2>&1
tail -40 $OUT/${TAG}_gputests.log
timeout 600 python tools/quick_bench.py sedov61 kh1000j kh2000j > $OUT/${TAG}_quick.log 2>&1
cat $OUT/${TAG}_quick.log
