#!/bin/bash
# scripts/build_variant.sh NAME "-DTG_LB=32 -DTG_PF=0": experiment build of libthrill_gpu.so into variants/NAME/ (git-ignored,
# travels with gpurun); select it with TG_LIB=variants/NAME/libthrill_gpu.so
set -e
cd "$(dirname "$0")/.."
name=$1; flags=$2
mkdir -p variants/$name
for f in thrill_b200/csrc/tg_*.cu; do
  b=$(basename $f .cu)
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v $flags -c $f -o variants/$name/$b.o 2> variants/$name/$b.ptxas.log &
done
wait
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o variants/$name/libthrill_gpu.so variants/$name/*.o -lnccl -lcudart
ls -la variants/$name/libthrill_gpu.so
