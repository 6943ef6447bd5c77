#!/bin/bash
mkdir -p gpurun_out
run() { BEVR_LIB=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-train --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', 'step %.3f fwd %.3f bwd %.3f' % (d['ms_per_step'], d['roofline']['fwd']['ms'], d['roofline']['bwd']['ms']))"; }
for v in v1 v2 v3 v4 v1; do run $v scripts/_build/libbevr_$v.so | tee -a gpurun_out/r2g_sweep.txt; done
