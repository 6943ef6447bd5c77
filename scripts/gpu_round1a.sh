#!/bin/bash
# first GPU pass: parity tests, smoke, bench (fp32 + bf16), tile sweep
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
python bench.py --steps 10 --warmup 3 --dtype bf16 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
for t in 8,8,256 8,8,512 16,8,256 16,16,256 16,16,512 16,16,1024 32,8,512 32,16,1024 32,32,1024 64,4,512; do
  echo "tile $t" >> gpurun_out/sweep.txt
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --tile $t >> gpurun_out/sweep.txt 2>&1
done
tail -5 gpurun_out/pytest.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench_fp32.json
