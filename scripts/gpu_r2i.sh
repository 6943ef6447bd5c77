#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rpe_attn_gpu.py tests/test_reference_gpu.py tests/test_modules.py tests/test_encoder_layer.py -m gpu -q --tb=short -s > gpurun_out/r2i_tests.log 2>&1
grep -E "passed|failed|FAILED|BEVRender param|^E  " gpurun_out/r2i_tests.log | head -30
python - <<'PY' 2>&1 | tail -5
import json, sys
sys.argv = ['bench.py']
import bench
print(json.dumps(bench.k3_leg()))
PY
N=200 C=256 H=8 D=4 B=1 ITERS=3 timeout 600 python scripts/bench_modules.py 2>&1 | tail -1
timeout 600 python scripts/bench_modules.py 2>&1 | tail -1
