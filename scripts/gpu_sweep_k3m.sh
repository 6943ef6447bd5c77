#!/bin/bash
# K3 mma.sync forward: resident CTAs x keys per softmax step
mkdir -p gpurun_out; rm -f gpurun_out/sweep_k3m.txt
cd bevrender_b200/csrc
for cfg in "4 32" "3 32" "5 32" "4 64" "3 64" "4 16" "2 64"; do
  set -- $cfg
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I../../include --expt-relaxed-constexpr \
    -DK3_MMA_MINB=$1 -DK3_MMA_CH=$2 -c rpe_attn.cu -o build/rpe_attn.o 2>/dev/null || { echo "cfg $cfg: compile failed" >> ../../gpurun_out/sweep_k3m.txt; continue; }
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../libbevr.so build/*.o -lcudart
  cd ../..
  echo "cfg minb=$1 CH=$2" >> gpurun_out/sweep_k3m.txt
  python scripts/exp_k3_atomics.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   tsa fwd %.3f  sca fwd %.3f'%(d['tsa_fwd'], d['sca_fwd']))" >> gpurun_out/sweep_k3m.txt
  cd bevrender_b200/csrc
done
cd ../..
cat gpurun_out/sweep_k3m.txt
