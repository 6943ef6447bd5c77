#!/bin/bash
# K3 key kernel: channel split across lanes (K3_CS) sweep
mkdir -p gpurun_out; rm -f gpurun_out/sweep_k3d.txt
cd bevrender_b200/csrc
for cfg in "1 2" "2 2" "2 3"; do
  set -- $cfg
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I../../include --expt-relaxed-constexpr \
    -DK3_CS=$1 -DK3_DKV_MINB2=$2 -c rpe_attn.cu -o build/rpe_attn.o 2>/dev/null || { echo "cfg $cfg: compile failed" >> ../../gpurun_out/sweep_k3d.txt; continue; }
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../libbevr.so build/*.o -lcudart
  cd ../..
  echo "cfg CS=$1 minb=$2" >> gpurun_out/sweep_k3d.txt
  python scripts/exp_k3_atomics.py 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('   tsa bwd %.3f (no table grad %.3f)   sca bwd %.3f (%.3f)'%(d['tsa_bwd_rpegrad1_posgrad1'], d['tsa_bwd_rpegrad0_posgrad0'], d['sca_bwd_rpegrad1_posgrad1'], d['sca_bwd_rpegrad0_posgrad0']))" >> gpurun_out/sweep_k3d.txt
  python scripts/bench_modules.py 2>/dev/null | cut -c40-200 >> gpurun_out/sweep_k3d.txt
  [ "$1" = "2" ] && python -m pytest tests/test_rpe_attn_gpu.py tests/test_modules.py -x -q -m gpu 2>&1 | tail -1 >> gpurun_out/sweep_k3d.txt
  cd bevrender_b200/csrc
done
cd ../..
cat gpurun_out/sweep_k3d.txt
