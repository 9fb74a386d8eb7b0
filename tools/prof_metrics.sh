#!/bin/bash
# a few memory-system counters for the fast-path kernels on the full-size workload (one replay pass each)
mkdir -p gpurun_out
W=${1:-C3}
CMD="python bench.py --workload $W --algo fast --steps 1 --warmup 3 --hxv-only ${2:-}"
$CMD > gpurun_out/plain3.log 2>&1 &&
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sector_hit_rate.pct,smsp__inst_executed.sum,sm__cycles_elapsed.max \
  --clock-control none -k regex:'k_fcol|k_srow|k_ccol' -s 6 -c 2 --csv --log-file gpurun_out/metrics_$W.csv $CMD > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log
cat gpurun_out/metrics_$W.csv | awk -F'","' 'NR>1{print $5"|"$(NF-2)"|"$(NF-1)"|"$NF}' | cut -c1-160
