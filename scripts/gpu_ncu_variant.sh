#!/bin/bash
# ncu --set full of the MSDA kernels of a variant library: gpu_ncu_variant.sh NAME [LIB]
mkdir -p gpurun_out
BEVR_LIB=$2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"msda_bwd_sorted" -c 1 -o gpurun_out/$1 -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-train --no-extra > gpurun_out/$1.log 2>&1; echo "ncu rc=$?"
