#!/bin/bash
# end-of-round re-measurement after the K3 mma work (1 GPU)
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/pytest_final.log 2>&1; tail -2 gpurun_out/pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; tail -1 gpurun_out/smoke_final.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
python scripts/prof_modules.py > gpurun_out/plain_rpe.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"rpe_attn|offset_head|rpe_combine|rpe_presplit" -c 16 -o gpurun_out/prof_rpe_r1c -f python scripts/prof_modules.py > gpurun_out/ncu_rpe.log 2>&1
python scripts/bench_modules.py > gpurun_out/modules.json 2>/dev/null
N=200 C=256 H=8 D=4 B=1 ITERS=2 python scripts/bench_modules.py > gpurun_out/modules_n200.json 2>/dev/null
python scripts/exp_k3_atomics.py > gpurun_out/k3_final.json 2>/dev/null
python bench.py --workload train --bev 56 --steps 3 --warmup 3 > gpurun_out/train_bev56_1gpu.json 2> gpurun_out/train56.err
python bench.py --workload train --bev 28 --steps 5 --warmup 3 > gpurun_out/train_bev28_1gpu.json 2> gpurun_out/train28.err
for f in bench_fp32 modules modules_n200 k3_final train_bev56_1gpu train_bev28_1gpu; do echo "== $f"; cut -c1-420 gpurun_out/$f.json; done
tail -1 gpurun_out/ncu_rpe.log
