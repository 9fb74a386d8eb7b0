#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "fast_hxv" 2>&1 | grep -E "Error|error|passed|failed|FAILED" | head -12
run() { name=$1; shift; timeout 300 python bench.py --steps 10 --warmup 3 --hxv-only "$@" > gpurun_out/hxv_$name.json 2> gpurun_out/hxv_$name.err; echo "$name: $(cut -c40-250 gpurun_out/hxv_$name.json)"; tail -2 gpurun_out/hxv_$name.err; }
run C3_cmax288 --workload C3 --algo fast --opt srow_cmax=288
run C3_cmax224 --workload C3 --algo fast --opt srow_cmax=224
run C3_cmax160 --workload C3 --algo fast --opt srow_cmax=160
run C3_lr4_224 --workload C3 --algo fast --opt srow_cmax=224 --opt srow_lr=4
