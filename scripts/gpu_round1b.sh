#!/bin/bash
# second GPU pass: all parity tests (K1-K4 + modules), module timings, ncu of the MSDA kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
python scripts/bench_modules.py > gpurun_out/modules.json 2> gpurun_out/modules.err
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:msda -s 6 -c 2 -o gpurun_out/prof_msda_r1b -f $CMD > gpurun_out/ncu_full.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1b.csv $CMD > gpurun_out/ncu_list.log 2>&1
tail -15 gpurun_out/pytest.log; cat gpurun_out/modules.json; tail -3 gpurun_out/modules.err; tail -3 gpurun_out/ncu_full.log
