#!/bin/bash
# K3 timing of the default library and of variant libraries (scripts/_build/libbevr_NAME.so): bench.k3_leg + parity
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rpe_attn_gpu.py -x -q -m gpu 2>&1 | tail -2
run() { BEVR_LIB=$2 timeout 600 python - "$1" <<'PY' | tee -a gpurun_out/k3_variants.txt
import sys, json
sys.path.insert(0, '.')
import bench
r = bench.k3_leg()
print(sys.argv[1], 'tsa fwd %.3f bwd %.3f | sca fwd %.3f bwd %.3f' % (r['tsa']['fwd_ms'], r['tsa']['bwd_ms'], r['sca']['fwd_ms'], r['sca']['bwd_ms']))
PY
}
run default ""
for v in "$@"; do run $v scripts/_build/libbevr_$v.so; done
