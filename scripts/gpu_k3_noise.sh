#!/bin/bash
run() { BEVR_LIB=$2 timeout 600 python - "$1" <<'PY'
import sys, json
sys.path.insert(0, '.')
import bench
r = bench.k3_leg()
print(sys.argv[1], 'tsa fwd %.3f bwd %.3f | sca fwd %.3f bwd %.3f' % (r['tsa']['fwd_ms'], r['tsa']['bwd_ms'], r['sca']['fwd_ms'], r['sca']['bwd_ms']))
PY
}
for i in 1 2 3; do run default ""; run fq4 scripts/_build/libbevr_fq4.so; run kvreg scripts/_build/libbevr_kvreg.so; done
