#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_msda_gpu.py -m gpu -q --tb=short > gpurun_out/r2f_tests.log 2>&1
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2f_tests.log | head -20
run() { BEVR_LIB=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', 'step %.3f fwd %.3f bwd %.3f' % (d['ms_per_step'], d['roofline']['fwd']['ms'], d['roofline']['bwd']['ms']))"; }
run default "" | tee -a gpurun_out/r2f_sweep.txt
for v in nv4_np1024 nv8_np2048 nv4_np2048 nv8_np512; do run $v scripts/_build/libbevr_$v.so | tee -a gpurun_out/r2f_sweep.txt; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_bwd_sorted -c 1 -o gpurun_out/r2f_bwd_sorted python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2f_ncu.log 2>&1
