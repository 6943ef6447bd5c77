#!/bin/bash
# build_variant.sh NAME "-DFOO=1 -DBAR=2": a variant libbevr (same ABI) in scripts/_build/libbevr_NAME.so for tuning sweeps
set -e
cd "$(dirname "$0")/../bevrender_b200/csrc"
out=../../scripts/_build; mkdir -p $out/obj_$1
for f in api msda project rpe_attn correlate warp offset_head recall; do
  if [ "$f" = "${3:-msda}" ] || [ ! -f build/$f.o ]; then
    nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -I../../include --expt-relaxed-constexpr $2 -c $f.cu -o $out/obj_$1/$f.o
  else cp build/$f.o $out/obj_$1/$f.o; fi
done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $out/libbevr_$1.so $out/obj_$1/*.o -lcudart
echo built $out/libbevr_$1.so
