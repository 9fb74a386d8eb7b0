#!/bin/bash
# ncu capture of the two fast-path kernels on a workload small enough for replay (default C2)
mkdir -p gpurun_out
W=${1:-C2}
CMD="python bench.py --workload $W --algo fast --steps 1 --warmup 3 --hxv-only"
$CMD > gpurun_out/plain2.log 2>&1 &&
timeout 420 ncu --set full --clock-control none --import-source on -k regex:'k_fcol|k_srow' -s 6 -c 2 -o gpurun_out/prof_fast_$W -f $CMD > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
ls -la gpurun_out/*.ncu-rep
