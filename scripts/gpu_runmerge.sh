#!/bin/bash
timeout 900 python -m pytest tests/test_rpe_attn_gpu.py tests/test_modules.py tests/test_reference_gpu.py tests/test_abi.py -q -m gpu 2>&1 | tail -2
python scripts/bench_modules.py 2>/dev/null | tail -1 | cut -c1-300
bash scripts/gpu_k3_variants.sh 2>&1 | tail -1
python bench.py --workload train --steps 6 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('train %.2f samples/s %.1f ms' % (d['value'], d['ms_per_step']))"
