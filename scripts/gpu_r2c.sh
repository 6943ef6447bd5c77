#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_msda_gpu.py -m gpu -q --tb=short -x > gpurun_out/r2c_msda.log 2>&1
tail -15 gpurun_out/r2c_msda.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
cat gpurun_out/r2c_bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['fwd'], d['roofline']['bwd'])"
timeout 900 python -m pytest tests/test_reference_gpu.py tests/test_rpe_attn_gpu.py -m gpu -q --tb=short -s > gpurun_out/r2c_ref.log 2>&1
tail -40 gpurun_out/r2c_ref.log
