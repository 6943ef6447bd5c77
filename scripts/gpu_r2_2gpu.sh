#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo "rc=$?"
tail -3 gpurun_out/r2_bench_2gpu.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_2gpu.json').read().strip().splitlines()[-1])
print('value', d['value'], 'n', d['n_gpus'], 'e2e', d['e2e']); print('train', json.dumps(d['train'])[:600])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_ref_2gpu.json 2>/dev/null; echo "ref rc=$?"; wc -l gpurun_out/r2_ref_2gpu.json
