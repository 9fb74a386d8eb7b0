#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 300 python bench.py --steps 10 --warmup 3 --hxv-only --workload C3 "$@" > gpurun_out/var_$name.json 2> gpurun_out/var_$name.err; echo "$name: $(cut -c40-250 gpurun_out/var_$name.json)"; tail -1 gpurun_out/var_$name.err; }
run base
for v in far4 far2 in2 far4in2; do EDGPU_LIB=dmft-lanc-ed_b200/variants/libedgpu_$v.so run $v; done
run base2
run cmax192 --opt srow_cmax=192
run cmax128 --opt srow_cmax=128
