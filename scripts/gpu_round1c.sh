#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
python bench.py --steps 10 --warmup 3 --dtype bf16 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
rm -f gpurun_out/sweep.txt
for t in 8,8,256 16,8,256 16,16,256 32,8,256 8,16,256 4,8,256; do
  echo "tile $t" >> gpurun_out/sweep.txt
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --tile $t >> gpurun_out/sweep.txt 2>&1
done
python scripts/prof_modules.py > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rpe -c 8 -o gpurun_out/prof_rpe_r1c -f python scripts/prof_modules.py > gpurun_out/ncu_rpe.log 2>&1
tail -12 gpurun_out/pytest.log; cat gpurun_out/bench_fp32.json; tail -2 gpurun_out/ncu_rpe.log
