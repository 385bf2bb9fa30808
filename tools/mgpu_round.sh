#!/bin/bash
# multi-GPU visit (gpurun --gpus N): bitwise slab tests + weak-scaling bench lines.  usage: bash tools/mgpu_round.sh <tag> <N...>
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -x 2>&1 | tail -15 > $OUT/${TAG}_mgpu_tests.log
cat $OUT/${TAG}_mgpu_tests.log
# each argument: N or N:workload (e.g. 8:sedov256)
for ARG in "$@"; do
  N=${ARG%%:*}; WL=""; SFX=""
  if [ "$ARG" != "$N" ]; then WL="--workload ${ARG#*:}"; SFX="_${ARG#*:}"; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 $WL > $OUT/${TAG}_bench_${N}gpu$SFX.json 2> $OUT/${TAG}_bench_${N}gpu$SFX.err
  tail -3 $OUT/${TAG}_bench_${N}gpu$SFX.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_bench_${N}gpu$SFX.json"))
    print("N=$N", d["config"]["workload"], "ms/step %.3f value %.3e e2e %s" % (d["ms_per_step"], d["value"], d["e2e"] and "%.3e" % d["e2e"]["value"]))
    print("   ", {k: round(v["ms_per_step"],4) for k,v in d["kernels"].items()})
    print("   roofline", d["roofline"])
except Exception as e:
    print("N=$N failed", e)
PY
done
