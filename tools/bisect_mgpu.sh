for so in libmlh_da57558.so libmlh_a8d9b80.so libmlh_gpu.so; do
  echo "=== $so"
  MLH_GPU_LIB=$PWD/meshlesshydro_b200/$so python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/mgpu_worker.py sedov 3 2>&1 | grep -E "MISMATCH|case=" | head -12
done
