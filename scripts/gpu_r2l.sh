#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_msda_gpu.py -m gpu -q --tb=short -k "bf16 or accumulate" > gpurun_out/r2l_tests.log 2>&1
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2l_tests.log | head
run() { BEVR_LIB=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-train --no-extra $3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', 'step %.3f fwd %.3f bwd %.3f' % (d['ms_per_step'], d['roofline']['fwd']['ms'], d['roofline']['bwd']['ms']))"; }
run default "" | tee -a gpurun_out/r2l_sweep.txt
for v in un2 un4 mb2048; do run $v scripts/_build/libbevr_$v.so | tee -a gpurun_out/r2l_sweep.txt; done
run bf16_f32grad "" "--dtype bf16" | tee -a gpurun_out/r2l_sweep.txt
run bf16_bf16grad "" "--dtype bf16 --bf16-out-grad" | tee -a gpurun_out/r2l_sweep.txt
