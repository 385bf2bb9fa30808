#!/bin/bash
# Multi-GPU visit on the final build: bitwise tests, the default bench line, the weak Sedov line, and the two ways out of a
# one-rank failure (exception / hang on rank 1 while rank 0 waits in NCCL).  usage: bash tools/visit_mgpu_final.sh <tag> <ngpus>
TAG=${1:-m}; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q --durations=8 > $OUT/${TAG}_mgpu_tests.log 2>&1
tail -15 $OUT/${TAG}_mgpu_tests.log
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 > $OUT/${TAG}_bench_${N}gpu.json 2> $OUT/${TAG}_bench_${N}gpu.err
echo "default bench exit $?"; cut -c1-600 $OUT/${TAG}_bench_${N}gpu.json
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 3 --workload sedov61 > $OUT/${TAG}_bench_${N}gpu_sedov61.json 2> $OUT/${TAG}_bench_${N}gpu_sedov61.err
echo "sedov61 weak exit $?"; cut -c1-400 $OUT/${TAG}_bench_${N}gpu_sedov61.json
for kind in raise hang; do
  MLH_BENCH_FAULT=mgpu_check:$kind:1 MLH_BENCH_SECTION_LIMIT=30 timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 \
      --workload kh1000 --min-seconds 0.2 > $OUT/${TAG}_fault_$kind.json 2> $OUT/${TAG}_fault_$kind.err
  echo "fault $kind exit $?"; python - <<PY
import json
try:
    l = json.loads(open("$OUT/${TAG}_fault_$kind.json").read().strip().splitlines()[-1])
    print("  line printed: value %.4g, mgpu_check = %s" % (l["value"], l["mgpu_check"]))
except Exception as e:
    print("  NO LINE:", e)
PY
done
