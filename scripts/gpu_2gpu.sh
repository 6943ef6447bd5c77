#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_fp32_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -3 gpurun_out/bench_2gpu.err
$TR bench.py --gpus 2 --workload train --bev 28 --steps 5 --warmup 3 > gpurun_out/train_bev28_2gpu.json 2> gpurun_out/train28_2.err; tail -3 gpurun_out/train28_2.err
$TR bench.py --gpus 2 --workload train --bev 56 --steps 3 --warmup 3 > gpurun_out/train_bev56_2gpu.json 2> gpurun_out/train56_2.err; tail -3 gpurun_out/train56_2.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
cat gpurun_out/bench_fp32_2gpu.json gpurun_out/train_bev28_2gpu.json gpurun_out/train_bev56_2gpu.json gpurun_out/bench_reference.json
