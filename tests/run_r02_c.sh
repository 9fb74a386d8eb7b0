#!/bin/bash
# GPU call C of round 2 (1 GPU): full GPU suite after the halo-axpy / window pipeline restructure + observables, bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 600 --timeout-method=thread 2>&1 | tail -40 > gpurun_out/r02c_pytest.log; tail -12 gpurun_out/r02c_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02c_smoke.log 2>&1; tail -2 gpurun_out/r02c_smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err; echo "bench1 rc=$?"; cut -c1-300 gpurun_out/r02c_bench_n1.json; tail -3 gpurun_out/r02c_bench_n1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02c_bench_ref.json 2> gpurun_out/r02c_bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r02c_bench_ref.json
