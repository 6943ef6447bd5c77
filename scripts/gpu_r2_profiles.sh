#!/bin/bash
# round-2 profile set on one B200: bench line, launch list, ncu --set full of the MSDA kernels, config-3 timings
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_fp32_1gpu.json 2> gpurun_out/r2_bench_fp32_1gpu.err; echo "bench rc=$?"
timeout 600 python bench.py --dtype bf16 --steps 20 --warmup 5 --no-train --no-extra > gpurun_out/r2_bench_bf16_1gpu.json 2>/dev/null; echo "bf16 rc=$?"
timeout 600 python scripts/bench_config3.py > gpurun_out/r2_config3.json 2> gpurun_out/r2_config3.err; echo "config3 rc=$?"; tail -2 gpurun_out/r2_config3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_msda.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-train --no-extra > /dev/null 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"msda_(fwd|bwd_sorted)" -c 2 -o gpurun_out/r2_msda python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-train --no-extra > gpurun_out/r2_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r2_*
