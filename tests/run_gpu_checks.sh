#!/bin/bash
# One GPU-box call: the whole GPU parity suite, smoke and the default bench line (see tools/prof_*.sh for ncu).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_C3.json 2> gpurun_out/bench_C3.err; cut -c1-200 gpurun_out/bench_C3.json; tail -3 gpurun_out/bench_C3.err
