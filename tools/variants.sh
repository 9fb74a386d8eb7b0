#!/bin/bash
# Build alternative libedgpu variants (compile-time kernel parameters) next to the product library:
#   tools/variants.sh name "-DSROW_NB_FAR=3 -DSROW_NB_IN=2" ...
# and time them on the GPU box with EDGPU_LIB=dmft-lanc-ed_b200/variants/libedgpu_<name>.so python bench.py ...
set -e
cd "$(dirname "$0")/../dmft-lanc-ed_b200/csrc"
mkdir -p ../variants build_var
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O3 --expt-relaxed-constexpr -I/usr/include"
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  $NVCC $FLAGS $defs -c hxv_fast.cu -o build_var/hxv_fast_$name.o
  $NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libedgpu_$name.so $(ls build/*.o | grep -v -e hxv_fast.o -e selftest.o) build_var/hxv_fast_$name.o -ldl
  echo built $name
done
