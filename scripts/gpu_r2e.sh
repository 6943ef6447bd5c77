#!/bin/bash
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/r2e_grid.log 2>&1
import torch
for n in (28, 50, 56, 100, 200, 250, 280):
    c = torch.arange(0, n, dtype=torch.float32).div(n - 1.0).mul(2.0).sub(1.0)
    g = torch.arange(0, n, dtype=torch.float32, device="cuda").div(n - 1.0).mul(2.0).sub(1.0).cpu()
    print(n, "CPU-vs-CUDA normalized grid mismatches:", int((c != g).sum()))
PY
cat gpurun_out/r2e_grid.log
timeout 900 python -m pytest tests/test_msda_gpu.py tests/test_reference_gpu.py tests/test_rpe_attn_gpu.py -m gpu -q --tb=short -s > gpurun_out/r2e_tests.log 2>&1
grep -E "passed|failed|FAILED|BEVRender param|^E  " gpurun_out/r2e_tests.log | head -40
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
python -c "import sys,json; d=json.load(open('gpurun_out/r2e_bench.json')); print(d['ms_per_step'], d['roofline']['fwd'], d['roofline']['bwd'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:msda_bwd_sorted -c 1 -o gpurun_out/r2e_bwd_sorted python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2e_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
