#!/bin/bash
# last occupancy check of the mma kernels after the pre-split staging (one change per config)
mkdir -p gpurun_out; rm -f gpurun_out/sweep_k3p.txt
cd bevrender_b200/csrc
for cfg in "5 3 3" "4 4 3" "4 3 2"; do
  set -- $cfg
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I../../include --expt-relaxed-constexpr \
    -DK3_MMA_MINB=$1 -DK3_DQ_MMA_MINB=$2 -DK3_DKV_MMA_MINB=$3 -c rpe_attn.cu -o build/rpe_attn.o 2>/dev/null || { echo "cfg $cfg: compile failed" >> ../../gpurun_out/sweep_k3p.txt; continue; }
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../libbevr.so build/*.o -lcudart
  cd ../..
  echo "cfg fwd_minb=$1 dq_minb=$2 dkv_minb=$3" >> gpurun_out/sweep_k3p.txt
  python scripts/exp_k3_atomics.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   tsa fwd %.3f bwd %.3f   sca fwd %.3f bwd %.3f'%(d['tsa_fwd'], d['tsa_bwd_rpegrad1_posgrad1'], d['sca_fwd'], d['sca_bwd_rpegrad1_posgrad1']))" >> gpurun_out/sweep_k3p.txt
  cd bevrender_b200/csrc
done
cd ../..
cat gpurun_out/sweep_k3p.txt
