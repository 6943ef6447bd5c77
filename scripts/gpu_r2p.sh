#!/bin/bash
mkdir -p gpurun_out
run() { BEVR_LIB=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-train --no-extra $3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', 'step %.3f fwd %.3f bwd %.3f' % (d['ms_per_step'], d['roofline']['fwd']['ms'], d['roofline']['bwd']['ms']))"; }
run default "" | tee -a gpurun_out/r2p_sweep.txt
for v in "$@"; do run $v scripts/_build/libbevr_$v.so | tee -a gpurun_out/r2p_sweep.txt; run ${v}_bf16 scripts/_build/libbevr_$v.so "--dtype bf16" | tee -a gpurun_out/r2p_sweep.txt; done
