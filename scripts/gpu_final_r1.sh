#!/bin/bash
# Round-1 final measurements on one B200: tests, contract bench lines, launch list, ncu captures, modules, training step
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/pytest_final.log 2>&1; tail -2 gpurun_out/pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -1 gpurun_out/smoke_final.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
python bench.py --steps 20 --warmup 5 --dtype bf16 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_msda.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_msda_r1.csv $CMD > gpurun_out/ncu_list.log 2>&1
python scripts/prof_modules.py > gpurun_out/plain_rpe.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"rpe_attn|offset_head|rpe_combine" -c 12 -o gpurun_out/prof_rpe_r1b -f python scripts/prof_modules.py > gpurun_out/ncu_rpe.log 2>&1
python scripts/bench_modules.py > gpurun_out/modules.json 2>/dev/null
N=200 C=256 H=8 D=4 B=1 ITERS=2 python scripts/bench_modules.py > gpurun_out/modules_n200.json 2>/dev/null
python scripts/bench_correlate.py > gpurun_out/correlate.json 2>/dev/null
python bench.py --workload train --bev 56 --steps 3 --warmup 3 > gpurun_out/train_bev56_1gpu.json 2> gpurun_out/train56.err
python bench.py --workload train --bev 28 --steps 5 --warmup 3 > gpurun_out/train_bev28_1gpu.json 2> gpurun_out/train28.err
for f in bench_fp32 bench_bf16 bench_reference modules modules_n200 correlate train_bev56_1gpu train_bev28_1gpu; do echo "== $f"; cut -c1-600 gpurun_out/$f.json; done
tail -1 gpurun_out/ncu_rpe.log
