#!/bin/bash
# K3 forward: rows per thread (R) x resident CTAs sweep; TSA / SCA shapes of exp_k3_atomics.py
mkdir -p gpurun_out; rm -f gpurun_out/sweep_k3r.txt
cd bevrender_b200/csrc
for cfg in "1 4 8" "2 4 8" "2 5 8" "2 6 8" "2 4 4" "2 5 4"; do
  set -- $cfg
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I../../include --expt-relaxed-constexpr \
    -DK3_R=$1 -DK3_R2_MINB=$2 -DK3_CH=$3 -c rpe_attn.cu -o build/rpe_attn.o 2>/dev/null || { echo "cfg $cfg: compile failed" >> ../../gpurun_out/sweep_k3r.txt; continue; }
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../libbevr.so build/*.o -lcudart
  cd ../..
  echo "cfg R=$1 minb=$2 CH=$3" >> gpurun_out/sweep_k3r.txt
  python scripts/exp_k3_atomics.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   tsa fwd %.3f  sca fwd %.3f   (bwd %.3f / %.3f)'%(d['tsa_fwd'], d['sca_fwd'], d['tsa_bwd_rpegrad1_posgrad1'], d['sca_bwd_rpegrad1_posgrad1']))" >> gpurun_out/sweep_k3r.txt
  cd bevrender_b200/csrc
done
cd ../..
python -m pytest tests/test_rpe_attn_gpu.py -x -q -m gpu 2>&1 | tail -2 >> gpurun_out/sweep_k3r.txt
cat gpurun_out/sweep_k3r.txt
