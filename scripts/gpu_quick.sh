#!/bin/bash
# quick GPU pass: parity tests + module timings + short bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
python scripts/bench_modules.py > gpurun_out/modules.json 2> gpurun_out/modules.err
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
grep -E "^E  .*(Error|assert)|FAILED|passed|failed" gpurun_out/pytest.log | head -30; cat gpurun_out/modules.json; tail -2 gpurun_out/modules.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1]); r = d['roofline']
print('step %.3f fwd %.3f bwd %.3f frac %.4f' % (d['ms_per_step'], r['fwd']['ms'], r['bwd']['ms'], r['fwd_bwd_frac']))
PY
