#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
$TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_fp32_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; tail -2 gpurun_out/bench_${N}gpu.err
$TR bench.py --gpus $N --workload train --bev 56 --steps 3 --warmup 3 > gpurun_out/train_bev56_${N}gpu.json 2> gpurun_out/train56_${N}.err; tail -2 gpurun_out/train56_${N}.err
python - <<PY
import json
for f in ("gpurun_out/bench_fp32_${N}gpu.json", "gpurun_out/train_bev56_${N}gpu.json"):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], d.get("e2e", {}) and d["e2e"].get("ms_per_step"), d["clocks"])
PY
