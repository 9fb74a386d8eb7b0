#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "fast_hxv or golden or lanc" 2>&1 | grep -E "^E  .*(rror|assert)|passed|failed|FAILED" | head -8
run() { name=$1; shift; timeout 300 python bench.py --steps 10 --warmup 3 --hxv-only "$@" > gpurun_out/hxv_$name.json 2> gpurun_out/hxv_$name.err; echo "$name: $(cut -c40-250 gpurun_out/hxv_$name.json)"; tail -2 gpurun_out/hxv_$name.err; }
run C3_base --workload C3 --algo fast
run C2_base --workload C2 --algo fast
