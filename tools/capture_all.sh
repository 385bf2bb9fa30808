TAG=r01o; OUT=gpurun_out; mkdir -p $OUT
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $B > $OUT/${TAG}_ncu_launches.log 2>&1
for K in k_face_states k_face_iterate k_face_setup k_face_finish k_face_index k_gradient_limit k_neighbours k_density_matrix k_flux_sum_update; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o $OUT/${TAG}_prof_$K -f $B > $OUT/${TAG}_ncu_full_$K.log 2>&1
done
ls -la $OUT | tail -30
