#!/bin/bash
# One GPU-box call: the whole GPU parity suite, smoke, the default bench line, a launch list with DRAM bytes.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_C3.json 2> gpurun_out/bench_C3.err; cut -c1-200 gpurun_out/bench_C3.json; tail -3 gpurun_out/bench_C3.err
timeout 300 python bench.py --workload C2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C2.json 2> gpurun_out/bench_C2.err; cut -c1-200 gpurun_out/bench_C2.json
timeout 300 python bench.py --workload C3 --stored --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C3_stored.json 2> gpurun_out/bench_C3_stored.err; cut -c1-200 gpurun_out/bench_C3_stored.json
CMD="python bench.py --workload C3 --steps 2 --warmup 3 --hxv-only"
$CMD > gpurun_out/plain.log 2>&1 && timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_C3.csv $CMD > gpurun_out/ncu1.log 2>&1
tail -2 gpurun_out/ncu1.log
