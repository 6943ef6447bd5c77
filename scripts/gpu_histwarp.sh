#!/bin/bash
timeout 1200 python -m pytest tests/test_reference_gpu.py tests/test_history_warp.py tests/test_compat.py tests/test_harness_gpu.py -q -m gpu 2>&1 | tail -3
python bench.py --workload train --steps 6 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('train %.2f samples/s %.1f ms' % (d['value'], d['ms_per_step']))"
timeout 600 python scripts/prof_train_step.py 2>/dev/null | python -c "
import sys,json; d=json.load(sys.stdin)
for k in ('ms_per_step_events','gpu_busy_ms_per_step','kernel_launches_per_step','ours_ms_per_step'): print(k, d[k])
print(d['ours_by_kernel_ms'])"
