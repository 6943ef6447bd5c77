#!/bin/bash
# gather-ceiling microbenchmark + compile-time sweep of the prefetching MSDA kernels
mkdir -p gpurun_out; rm -f gpurun_out/sweep_msda2.txt
./scripts/_build/exp_gather_ceiling > gpurun_out/exp_gather_ceiling.txt 2>&1
cd bevrender_b200/csrc
for cfg in "6 4 4 2" "5 4 4 2" "4 3 4 2" "4 4 8 2" "5 3 4 1" "3 2 4 2"; do
  set -- $cfg
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I../../include --expt-relaxed-constexpr \
    -DMSDA_FWD_MINB=$1 -DMSDA_BWD_MINB=$2 -DMSDA_FWD_UNROLL=$3 -DMSDA_BWD_UNROLL=$4 -c msda.cu -o build/msda.o 2>/dev/null || { echo "cfg $cfg: compile failed" >> ../../gpurun_out/sweep_msda2.txt; continue; }
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../libbevr.so build/*.o -lcudart
  cd ../..
  echo "cfg fwd_minb=$1 bwd_minb=$2 fwd_unroll=$3 bwd_unroll=$4" >> gpurun_out/sweep_msda2.txt
  for dt in fp32 bf16; do
  python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --dtype $dt 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('   $dt step %.3f fwd %.3f bwd %.3f'%(d['ms_per_step'], r['fwd']['ms'], r['bwd']['ms']))" >> gpurun_out/sweep_msda2.txt
  done
  cd bevrender_b200/csrc
done
cd ../..
python -m pytest tests/test_msda_gpu.py -x -q -m gpu 2>&1 | tail -3 >> gpurun_out/sweep_msda2.txt
cat gpurun_out/exp_gather_ceiling.txt gpurun_out/sweep_msda2.txt
