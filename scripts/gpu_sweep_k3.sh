#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/sweep_k3.txt
cd bevrender_b200/csrc
for cfg in "128 16 1 1" "128 8 1 1" "128 8 4 1" "64 16 1 1" "64 8 1 1" "64 8 8 1" "64 16 6 1" "128 16 4 1" "128 16 1 3" "64 16 1 6" "256 16 1 1" "32 16 1 1"; do
  set -- $cfg
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I../../include --expt-relaxed-constexpr \
    -DK3_TQ=$1 -DK3_CH=$2 -DK3_FWD_MINB=$3 -DK3_DQ_MINB=$4 -Xptxas -v -c rpe_attn.cu -o build/rpe_attn.o 2> build/k3.log || { echo "cfg $cfg: compile failed" >> ../../gpurun_out/sweep_k3.txt; continue; }
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../libbevr.so build/*.o -lcudart
  regs=$(grep -A1 "rpe_attn_fwd_kernelILi32" build/k3.log | grep -o "Used [0-9]* registers" | head -1)
  regs2=$(grep -A1 "rpe_attn_bwd_dq_kernelILi32" build/k3.log | grep -o "Used [0-9]* registers" | head -1)
  cd ../..
  echo "cfg TQ=$1 CH=$2 fwd_minb=$3 dq_minb=$4 fwd:$regs dq:$regs2" >> gpurun_out/sweep_k3.txt
  python scripts/exp_k3_atomics.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   tsa fwd %.3f bwd %.3f | sca fwd %.3f bwd %.3f'%(d['tsa_fwd'], d['tsa_bwd_rpegrad1_posgrad1'], d['sca_fwd'], d['sca_bwd_rpegrad1_posgrad1']))" >> gpurun_out/sweep_k3.txt
  cd bevrender_b200/csrc
done
cat ../../gpurun_out/sweep_k3.txt
