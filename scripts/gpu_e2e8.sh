#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
lscpu | grep -E "NUMA|Socket|^CPU\(s\)|Model name" > gpurun_out/lscpu.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"
$TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_fp32_${N}gpu_numa.json 2> gpurun_out/bench_${N}gpu_numa.err; tail -2 gpurun_out/bench_${N}gpu_numa.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_fp32_${N}gpu_numa.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"])
PY
cat gpurun_out/lscpu.txt; head -12 gpurun_out/topo.txt | cut -c1-150
