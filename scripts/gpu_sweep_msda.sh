#!/bin/bash
# compile-time sweep of the MSDA kernels (occupancy hints / unroll); nvcc is available on the GPU box image
mkdir -p gpurun_out; rm -f gpurun_out/sweep_msda.txt
cd bevrender_b200/csrc
for cfg in "1 1 4 2" "6 1 4 2" "8 1 4 2" "1 1 8 2" "6 1 8 2" "1 1 16 2" "1 4 4 2" "1 5 4 2" "1 1 4 4" "1 4 4 4" "1 1 4 1"; do
  set -- $cfg
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I../../include --expt-relaxed-constexpr \
    -DMSDA_FWD_MINB=$1 -DMSDA_BWD_MINB=$2 -DMSDA_FWD_UNROLL=$3 -DMSDA_BWD_UNROLL=$4 -c msda.cu -o build/msda.o 2>/dev/null || { echo "cfg $cfg: compile failed" >> ../../gpurun_out/sweep_msda.txt; continue; }
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../libbevr.so build/*.o -lcudart
  cd ../..
  echo "cfg fwd_minb=$1 bwd_minb=$2 fwd_unroll=$3 bwd_unroll=$4" >> gpurun_out/sweep_msda.txt
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('   step %.3f fwd %.3f bwd %.3f'%(d['ms_per_step'], r['fwd']['ms'], r['bwd']['ms']))" >> gpurun_out/sweep_msda.txt
  cd bevrender_b200/csrc
done
cat ../../gpurun_out/sweep_msda.txt
