#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_modules.py > gpurun_out/plain_rpe.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rpe_attn -c 6 -o gpurun_out/prof_rpe_r1 -f python scripts/prof_modules.py > gpurun_out/ncu_rpe.log 2>&1
python scripts/prof_correlate.py > gpurun_out/plain_corr.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:corr_umma -c 1 -o gpurun_out/prof_corr_r1 -f python scripts/prof_correlate.py > gpurun_out/ncu_corr.log 2>&1
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_msda.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:msda -s 6 -c 2 -o gpurun_out/prof_msda_r1 -f $CMD > gpurun_out/ncu_msda.log 2>&1
$CMD > gpurun_out/plain_msda2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_msda_r1.csv $CMD > gpurun_out/ncu_list.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
python bench.py --steps 20 --warmup 5 --dtype bf16 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
python scripts/bench_correlate.py > gpurun_out/correlate.json 2>/dev/null
python scripts/bench_modules.py > gpurun_out/modules.json 2>/dev/null
for f in ncu_msda ncu_rpe ncu_corr; do tail -1 gpurun_out/$f.log; done
cat gpurun_out/bench_bf16.json | cut -c1-300
