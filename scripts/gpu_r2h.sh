#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; tail -3 gpurun_out/r2h_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r2h_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2h_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'], 'fwd', d['roofline']['fwd']['ms'], 'bwd', d['roofline']['bwd']['ms'])
print('e2e', d['e2e']); print('cpu', d['cpu_baseline'])
for k in ('train', 'k3', 'k5', 'config1'): print(k, json.dumps(d[k])[:700])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2h_ref.json 2>&1; tail -c 600 gpurun_out/r2h_ref.json
timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/r2h_tests.log 2>&1; tail -5 gpurun_out/r2h_tests.log
