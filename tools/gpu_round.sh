#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list + full capture of the flux kernel.
# usage (under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > $OUT/gpu_tests_$TAG.log
tail -3 $OUT/gpu_tests_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -c 3000 $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_launches_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_face_riemann -s 3 -c 1 -o $OUT/prof_k4b_$TAG -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full_$TAG.log 2>&1
timeout 600 ncu --clock-control none -k regex:k_face_riemann -s 3 -c 1 --csv --log-file $OUT/k4b_instmix_$TAG.csv \
    --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__cycles_elapsed.max,dram__bytes_read.sum,dram__bytes_write.sum \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_instmix_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_face_states -s 3 -c 1 -o $OUT/prof_k4a_$TAG -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full_k4a_$TAG.log 2>&1
ls -la $OUT | tail -20
