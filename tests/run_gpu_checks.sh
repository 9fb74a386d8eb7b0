#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "fast_hxv" 2>&1 | tail -3
run() { name=$1; shift; timeout 300 python bench.py --steps 10 --warmup 3 --hxv-only "$@" > gpurun_out/hxv_$name.json 2> gpurun_out/hxv_$name.err; echo "$name: $(cut -c40-140 gpurun_out/hxv_$name.json)"; tail -2 gpurun_out/hxv_$name.err; }
run C3_base --workload C3 --algo fast
run C3_cmax288 --workload C3 --algo fast --opt srow_cmax=288
run C3_cmax224 --workload C3 --algo fast --opt srow_cmax=224
run C3_cmax160 --workload C3 --algo fast --opt srow_cmax=160
run C3_lr4_288 --workload C3 --algo fast --opt srow_cmax=288 --opt srow_lr=4
