#!/bin/bash
# K3 backward: mma.sync kernels on/off and their occupancy
mkdir -p gpurun_out; rm -f gpurun_out/sweep_k3n.txt
cd bevrender_b200/csrc
for cfg in "1 3 0 2" "1 3 1 2" "1 3 1 3" "1 4 1 2" "1 2 1 2" "0 3 1 2"; do
  set -- $cfg
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I../../include --expt-relaxed-constexpr \
    -DK3_DQ_MMA=$1 -DK3_DQ_MMA_MINB=$2 -DK3_DKV_MMA=$3 -DK3_DKV_MMA_MINB=$4 -c rpe_attn.cu -o build/rpe_attn.o 2>/dev/null || { echo "cfg $cfg: compile failed" >> ../../gpurun_out/sweep_k3n.txt; continue; }
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../libbevr.so build/*.o -lcudart
  cd ../..
  echo "cfg dq_mma=$1 dq_minb=$2 dkv_mma=$3 dkv_minb=$4" >> gpurun_out/sweep_k3n.txt
  python scripts/exp_k3_atomics.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   tsa bwd %.3f (no table grad %.3f)   sca bwd %.3f (%.3f)'%(d['tsa_bwd_rpegrad1_posgrad1'], d['tsa_bwd_rpegrad0_posgrad0'], d['sca_bwd_rpegrad1_posgrad1'], d['sca_bwd_rpegrad0_posgrad0']))" >> gpurun_out/sweep_k3n.txt
  cd bevrender_b200/csrc
done
cd ../..
cat gpurun_out/sweep_k3n.txt
