#!/bin/bash
# One GPU-box visit with the final build of a round: parity tests, FP64-op / DRAM counts (-> profiles/fp64_ops.json),
# bench lines, ncu launch list and full captures of the main kernels.   usage (under gpurun): bash tools/gpu_round2.sh <tag> [quick]
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -5 > $OUT/${TAG}_gputests.log
tail -3 $OUT/${TAG}_gputests.log
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
for WL in sedov61 kh2000 kh1000; do
  B="python bench.py --workload $WL --steps 5 --warmup 3 --min-seconds 0 --no-cpu-baseline --no-e2e"
  # 3 warm-up steps and the restart (+1 step) come first (<= 100 launches); the window then covers timed steps, of which
  # tools/ncu_ops.py takes the last complete one that does not follow an upload
  timeout 900 ncu --clock-control none -k regex:'k_' -s 100 -c 90 --csv --log-file $OUT/${TAG}_ops_$WL.csv --metrics $M $B > $OUT/${TAG}_ncu_ops_$WL.log 2>&1
  python tools/ncu_ops.py $WL $OUT/${TAG}_ops_$WL.csv profiles/fp64_ops.json > $OUT/${TAG}_ops_$WL.txt; cat $OUT/${TAG}_ops_$WL.txt
done
cp profiles/fp64_ops.json $OUT/fp64_ops.json
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -5 $OUT/${TAG}_bench.err
timeout 600 python bench.py --workload sedov61 --steps 20 --warmup 3 > $OUT/${TAG}_bench_sedov61.json 2>> $OUT/${TAG}_bench.err
timeout 600 python bench.py --workload kh1000 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_kh1000.json 2>> $OUT/${TAG}_bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
python - <<PY
import json
for f in ("$OUT/${TAG}_bench.json", "$OUT/${TAG}_bench_sedov61.json", "$OUT/${TAG}_bench_kh1000.json"):
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f, "FAILED", e); continue
    print(f, "ms/step",d["ms_per_step"],"value %.3e"%d["value"],"e2e %.3e"%d["e2e"]["value"] if d.get("e2e") else None, d["timed_region"])
    print("  roofline", {k:d["roofline"].get(k) for k in ("kernel","bound","achieved","peak","frac","share_of_step","traffic")})
    for k,v in d["kernels"].items(): print("  %-24s %.4f  %s"%(k,v["ms_per_step"], d["kernel_rooflines"].get(k)))
print(open("$OUT/${TAG}_bench_reference.json").read()[:700])
PY
[ "$2" = "quick" ] && exit 0
B="python bench.py --steps 2 --warmup 3 --min-seconds 0 --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $B > $OUT/${TAG}_ncu_launches.log 2>&1
BS="python bench.py --workload sedov61 --steps 2 --warmup 3 --min-seconds 0 --no-cpu-baseline --no-e2e"
for K in k_face_states k_face_iterate k_face_setup k_face_finish k_face_index k_gradient_limit k_neighbours_cell k_density_matrix k_flux_sum_update; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 -o $OUT/${TAG}_prof_$K -f $BS > $OUT/${TAG}_ncu_full_$K.log 2>&1
  python tools/ncu_summary.py $OUT/${TAG}_prof_$K.ncu-rep > $OUT/${TAG}_${K}_ncu_full.txt 2>&1
done
# the dominant kernel of the headline workload (KH 4 M)
for K in k_face_iterate k_neighbours_cell; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 4 -c 1 -o $OUT/${TAG}_prof_kh2000_$K -f $B > $OUT/${TAG}_ncu_full_kh2000_$K.log 2>&1
  python tools/ncu_summary.py $OUT/${TAG}_prof_kh2000_$K.ncu-rep > $OUT/${TAG}_kh2000_${K}_ncu_full.txt 2>&1
done
python tools/ncu_lines.py $OUT/${TAG}_prof_kh2000_k_face_iterate.ncu-rep 40 > $OUT/${TAG}_kh2000_k_face_iterate_lines.txt 2>&1
ls -la $OUT | tail -40
