#!/bin/bash
# GPU call D of round 2 (1 GPU): cluster column kernel (entries resident in shared memory) -- parity, timings, counters
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "fast_hxv and C1" -p no:cacheprovider --timeout 100 --timeout-method=thread 2>&1 | tail -15 > gpurun_out/r02d_pytest_first.log; tail -4 gpurun_out/r02d_pytest_first.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "fast or sharded or lanczos or golden or gf_chains" -p no:cacheprovider --timeout 300 --timeout-method=thread 2>&1 | tail -60 > gpurun_out/r02d_pytest_fast.log; tail -8 gpurun_out/r02d_pytest_fast.log
run() { name=$1; shift; timeout 300 python bench.py --steps 10 --warmup 3 --hxv-only "$@" > gpurun_out/r02d_$name.json 2> gpurun_out/r02d_$name.err; echo "$name: $(cut -c40-260 gpurun_out/r02d_$name.json)"; tail -1 gpurun_out/r02d_$name.err; }
run c3 --workload C3
run c3_noccol --workload C3 --opt no_ccol=1
run c3v --workload C3V
run c3v_noccol --workload C3V --opt no_ccol=1
run c3_stored --workload C3 --stored
run c2 --workload C2
run c2_noccol --workload C2 --opt no_ccol=1
run c5 --workload C5 --steps 3
run c5_noccol --workload C5 --steps 3 --opt no_ccol=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "c5_ns18 or full_size" -p no:cacheprovider --timeout 500 --timeout-method=thread 2>&1 | tail -20 > gpurun_out/r02d_pytest_big.log; tail -4 gpurun_out/r02d_pytest_big.log
bash tools/prof_metrics.sh C3 > gpurun_out/r02d_metrics.log 2>&1; cp gpurun_out/metrics_C3.csv gpurun_out/r02d_metrics_C3.csv; tail -3 gpurun_out/r02d_metrics.log
