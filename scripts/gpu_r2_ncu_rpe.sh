#!/bin/bash
# round-2 ncu --set full capture of the K3 kernels (TSA then SCA, N=56, B=8, a third of the SCA reference points masked) + K5
mkdir -p gpurun_out
python scripts/prof_modules.py > gpurun_out/r2_plain_rpe.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rpe_attn_(fwd|bwd)" -c 6 -o gpurun_out/r2_rpe -f python scripts/prof_modules.py > gpurun_out/r2_ncu_rpe.log 2>&1; echo "ncu rpe rc=$?"
python scripts/prof_correlate.py > gpurun_out/r2_plain_corr.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:corr_umma -c 1 -o gpurun_out/r2_corr -f python scripts/prof_correlate.py > gpurun_out/r2_ncu_corr.log 2>&1; echo "ncu corr rc=$?"
