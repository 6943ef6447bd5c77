#!/bin/bash
# MSDA variants: parity of the FIRST variant library, then timing of the default and every variant
mkdir -p gpurun_out
BEVR_LIB=scripts/_build/libbevr_$1.so timeout 900 python -m pytest tests/test_msda_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -3
run() { BEVR_LIB=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-train --no-extra $3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', 'step %.3f fwd %.3f bwd %.3f' % (d['ms_per_step'], d['roofline']['fwd']['ms'], d['roofline']['bwd']['ms']))"; }
run default "" | tee -a gpurun_out/r2o_sweep.txt
for v in "$@"; do run $v scripts/_build/libbevr_$v.so | tee -a gpurun_out/r2o_sweep.txt; done
run default_bf16 "" "--dtype bf16" | tee -a gpurun_out/r2o_sweep.txt
for v in "$@"; do run ${v}_bf16 scripts/_build/libbevr_$v.so "--dtype bf16" | tee -a gpurun_out/r2o_sweep.txt; done
