#!/bin/bash
# contract run at 8 GPUs (driver's command line): value / e2e / train legs under torchrun
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo "rc=$?"
tail -3 gpurun_out/r2_bench_${N}gpu.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r2_bench_${N}gpu.json').read().strip().splitlines()[-1])
print('value', d['value'], 'n', d['n_gpus'], 'e2e', d['e2e']); print('train', json.dumps(d['train'])[:700])
PY
