#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/sweep_k3b.txt
cd bevrender_b200/csrc
for cfg in "64 1" "64 2" "64 3" "32 1" "32 3" "128 1" "128 2"; do
  set -- $cfg
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I../../include --expt-relaxed-constexpr \
    -DK3_TQB=$1 -DK3_DKV_MINB=$2 -c rpe_attn.cu -o build/rpe_attn.o 2> build/k3.log || { echo "cfg $cfg: compile failed" >> ../../gpurun_out/sweep_k3b.txt; continue; }
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../libbevr.so build/*.o -lcudart
  cd ../..
  echo "cfg TQB=$1 dkv_minb=$2" >> gpurun_out/sweep_k3b.txt
  python scripts/exp_k3_atomics.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   tsa fwd %.3f bwd %.3f | sca fwd %.3f bwd %.3f'%(d['tsa_fwd'], d['tsa_bwd_rpegrad1_posgrad1'], d['sca_fwd'], d['sca_bwd_rpegrad1_posgrad1']))" >> gpurun_out/sweep_k3b.txt
  cd bevrender_b200/csrc
done
cat ../../gpurun_out/sweep_k3b.txt
