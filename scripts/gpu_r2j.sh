#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_msda_gpu.py -m gpu -q --tb=short > gpurun_out/r2j_tests.log 2>&1
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2j_tests.log | head -20
run() { BEVR_LIB=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-train --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', 'step %.3f fwd %.3f bwd %.3f' % (d['ms_per_step'], d['roofline']['fwd']['ms'], d['roofline']['bwd']['ms']))"; }
run uniform_nv4_np1024 "" | tee -a gpurun_out/r2j_sweep.txt
for v in u_nv8 u_nv4_np2048 u_nv8_np2048 u_nv4_minb4; do run $v scripts/_build/libbevr_$v.so | tee -a gpurun_out/r2j_sweep.txt; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_bwd_sorted -c 1 -o gpurun_out/r2j_bwd_sorted python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-train --no-extra > gpurun_out/r2j_ncu.log 2>&1
