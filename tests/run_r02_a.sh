#!/bin/bash
# GPU call A of round 2: new row kernel (aligned chunks, group records) + sharded lists/halo (emulated), variants, C5, counters
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02a_smoke.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "fast or sharded" -p no:cacheprovider --timeout 300 --timeout-method=thread 2>&1 | tail -60 > gpurun_out/r02a_pytest_fast.log; tail -8 gpurun_out/r02a_pytest_fast.log
run() { name=$1; shift; timeout 300 python bench.py --steps 10 --warmup 3 --hxv-only --workload C3 "$@" > gpurun_out/r02a_$name.json 2> gpurun_out/r02a_$name.err; echo "$name: $(cut -c40-260 gpurun_out/r02a_$name.json)"; tail -1 gpurun_out/r02a_$name.err; }
run base
for v in nb2 nb3 nb4in1 st3; do EDGPU_LIB=dmft-lanc-ed_b200/variants/libedgpu_$v.so run $v; done
run t4 --opt srow_t=5
run t3 --opt srow_t=4
run stored --stored
timeout 900 python -m pytest tests -m gpu -q --tb=short -k "not fast and not sharded" -p no:cacheprovider --timeout 600 --timeout-method=thread 2>&1 | tail -40 > gpurun_out/r02a_pytest_rest.log; tail -5 gpurun_out/r02a_pytest_rest.log
timeout 500 python tools/c5_check.py > gpurun_out/r02a_c5.log 2>&1; tail -3 gpurun_out/r02a_c5.log
bash tools/prof_metrics.sh C3 > gpurun_out/r02a_metrics.log 2>&1; cp gpurun_out/metrics_C3.csv gpurun_out/r02a_metrics_C3.csv; tail -3 gpurun_out/r02a_metrics.log
