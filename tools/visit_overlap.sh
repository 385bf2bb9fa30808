#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
{
echo "=== default grids"; timeout 300 python tools/overlap_probe.py 2000
echo "=== half footprint (states 2, setup 4, finish 4, iterate 3 blocks/SM)"
MLH_GRID_STATES=2 MLH_GRID_SETUP=4 MLH_GRID_FINISH=4 MLH_GRID_ITERATE=3 timeout 300 python tools/overlap_probe.py 2000
echo "=== iterate 3 only"
MLH_GRID_ITERATE=3 timeout 300 python tools/overlap_probe.py 2000
echo "=== iterate 4, streaming kernels 3 waves of half footprint (states 6, setup 12, finish 12)"
MLH_GRID_STATES=6 MLH_GRID_SETUP=12 MLH_GRID_FINISH=12 MLH_GRID_ITERATE=4 timeout 300 python tools/overlap_probe.py 2000
} 2>&1 | tee $OUT/r2x_overlap.log
