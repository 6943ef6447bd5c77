#!/bin/bash
# sorted MSDA backward with next-batch prefetch / step 4 from shared memory: parity + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_msda_gpu.py tests/test_recall.py -m gpu -q --tb=short -x > gpurun_out/r2m_tests.log 2>&1
grep -E "passed|failed|FAILED|^E  " gpurun_out/r2m_tests.log | head
run() { BEVR_LIB=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-train --no-extra $3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', 'step %.3f fwd %.3f bwd %.3f' % (d['ms_per_step'], d['roofline']['fwd']['ms'], d['roofline']['bwd']['ms']))"; }
run default "" | tee -a gpurun_out/r2m_sweep.txt
for v in "$@"; do run $v scripts/_build/libbevr_$v.so | tee -a gpurun_out/r2m_sweep.txt; done
