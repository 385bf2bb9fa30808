#!/bin/bash
# A/B timing of alternative builds of the library: MLH_GPU_LIB=<so> quick_bench; "env:VAR=VAL,VAR=VAL" items set
# environment overrides (grid sizes) for the default library
for so in "$@"; do
  echo "=== $so"
  if [[ "$so" == env:* ]]; then
    env $(echo "${so#env:}" | tr ',' ' ') timeout 300 python tools/quick_bench.py sedov61 kh1000j 2>&1 | grep -E "N=|k0|k1|k4|k3|k2"
  else
    MLH_GPU_LIB=$PWD/meshlesshydro_b200/$so timeout 300 python tools/quick_bench.py sedov61 kh1000j 2>&1 | grep -E "N=|k0|k1|k4|k3|k2"
  fi
done
