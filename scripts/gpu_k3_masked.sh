#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $1" | tee -a gpurun_out/k3_masked.txt; BEVR_LIB=$2 timeout 600 python scripts/exp_k3_masked.py 2>/dev/null | tr -d '\n' | tee -a gpurun_out/k3_masked.txt; echo | tee -a gpurun_out/k3_masked.txt; }
run default ""
for v in "$@"; do run $v scripts/_build/libbevr_$v.so; done
