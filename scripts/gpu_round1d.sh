#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "^E  .*(Error|assert)|FAILED|passed|failed" gpurun_out/pytest.log | head -20
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; tail -3 gpurun_out/bench_fp32.err
python bench.py --workload train --bev 28 --steps 5 --warmup 3 > gpurun_out/train_bev28_1gpu.json 2> gpurun_out/train28.err; tail -3 gpurun_out/train28.err
python bench.py --workload train --bev 56 --steps 3 --warmup 3 > gpurun_out/train_bev56_1gpu.json 2> gpurun_out/train56.err; tail -3 gpurun_out/train56.err
cat gpurun_out/bench_fp32.json gpurun_out/train_bev28_1gpu.json gpurun_out/train_bev56_1gpu.json
