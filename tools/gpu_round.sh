#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list + full captures of the face kernels.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [quick]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > $OUT/${TAG}_gputests.log
tail -3 $OUT/${TAG}_gputests.log
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -5 $OUT/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("$OUT/${TAG}_bench.json"))
print("ms/step",d["ms_per_step"],"value",d["value"],"e2e",d["e2e"]["value"] if d.get("e2e") else None, "roofline", d["roofline"].get("frac"))
for k,v in d["kernels"].items(): print("  %-24s %.4f"%(k,v["ms_per_step"]))
PY
[ "$2" = "quick" ] && exit 0
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $B > $OUT/${TAG}_ncu_launches.log 2>&1
for K in k_face_riemann k_face_states k_face_index k_gradient_limit k_neighbours k_density_matrix k_flux_sum_update; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o $OUT/${TAG}_prof_$K -f $B > $OUT/${TAG}_ncu_full_$K.log 2>&1
done
timeout 600 ncu --clock-control none -k regex:'k_face_riemann|k_face_states' -s 6 -c 2 --csv --log-file $OUT/${TAG}_instmix.csv \
    --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__cycles_elapsed.max,dram__bytes_read.sum,dram__bytes_write.sum \
    $B > $OUT/${TAG}_ncu_instmix.log 2>&1
ls -la $OUT | tail -30
