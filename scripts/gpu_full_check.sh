#!/bin/bash
# what the driver runs at round end: the whole -m gpu suite, smoke(), the default bench line and the reference arm
mkdir -p gpurun_out
T=${1:-r2n}
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${T}_smoke.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_ref.json 2>/dev/null; echo "ref rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
print('value %.4g ms %.3f frac %.4f fwd %.3f bwd %.3f' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['fwd']['ms'], d['roofline']['bwd']['ms']))
print('e2e', d['e2e']); print('train', (d['train'] or {}).get('samples_per_s')); print('k3', json.dumps(d['k3'])[:400]); print('k5', d['k5']); print('c2v', d.get('config2_variants')); print('traffic', d['roofline']['traffic'])
PY
