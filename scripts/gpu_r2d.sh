#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_msda_gpu.py tests/test_reference_gpu.py tests/test_rpe_attn_gpu.py tests/test_modules.py tests/test_encoder_layer.py -m gpu -q --tb=short -s > gpurun_out/r2d_tests.log 2>&1
grep -E "passed|failed|FAILED|BEVRender param|^E  " gpurun_out/r2d_tests.log | head -40
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
python -c "import sys,json; d=json.load(open('gpurun_out/r2d_bench.json')); print(d['ms_per_step'], d['roofline']['fwd'], d['roofline']['bwd'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_bwd_sorted -c 1 -o gpurun_out/r2d_bwd_sorted python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2d_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
