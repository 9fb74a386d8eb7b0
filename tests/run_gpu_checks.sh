#!/bin/bash
# One GPU-box call: the whole GPU parity suite, smoke and the default bench line (see tools/prof_*.sh for ncu).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 600 --timeout-method=thread 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_C3.json 2> gpurun_out/bench_C3.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_C3.json; tail -3 gpurun_out/bench_C3.err
