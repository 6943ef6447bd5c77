#!/bin/bash
for v in 1 0; do echo "BEVR_SCA_SORT_KEYS=$v"; BEVR_SCA_SORT_KEYS=$v python scripts/bench_modules.py 2>/dev/null | tail -1 | cut -c1-300; done
timeout 900 python -m pytest tests/test_modules.py tests/test_reference_gpu.py tests/test_encoder_layer.py -q -m gpu -x 2>&1 | tail -3
for v in 1 0; do echo "BEVR_SCA_SORT_KEYS=$v"; BEVR_SCA_SORT_KEYS=$v python bench.py --workload train --steps 5 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('train %.2f samples/s %.1f ms' % (d['value'], d['ms_per_step']))"; done
