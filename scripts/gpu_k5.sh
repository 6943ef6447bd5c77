#!/bin/bash
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_correlate_gpu.py -x -q > gpurun_out/pytest_k5.log 2>&1; echo "k5 rc=$?" >> gpurun_out/pytest_k5.log
grep -E "^E  |FAILED|passed|failed|rc=" gpurun_out/pytest_k5.log | head -20
timeout 200 python scripts/bench_correlate.py > gpurun_out/correlate.json 2> gpurun_out/correlate.err; cat gpurun_out/correlate.json; tail -3 gpurun_out/correlate.err
