#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_msda_gpu.py -m gpu -q --tb=short > gpurun_out/r2k_tests.log 2>&1
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2k_tests.log | head -20
for i in 1 2; do timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-train --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fp32 step %.3f fwd %.3f bwd %.3f frac %.4f' % (d['ms_per_step'], d['roofline']['fwd']['ms'], d['roofline']['bwd']['ms'], d['roofline']['frac']))"; done
timeout 300 python bench.py --dtype bf16 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-train --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bf16 step %.3f fwd %.3f bwd %.3f frac %.4f' % (d['ms_per_step'], d['roofline']['fwd']['ms'], d['roofline']['bwd']['ms'], d['roofline']['frac']))"
