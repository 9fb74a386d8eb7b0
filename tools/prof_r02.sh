#!/bin/bash
# Round-2 ncu evidence for the final build, one B200 (each capture only after the plain command exited 0):
#  1. launch list of a short C3 H*v loop: device time + DRAM bytes per launch
#  2. memory-system counters of k_srow / k_fcol on C3 (single pass)
#  3. --set full --import-source on capture of both kernels on C2 (replays fit one call)
mkdir -p gpurun_out
CMD3="python bench.py --workload C3 --algo fast --steps 2 --warmup 3 --hxv-only"
$CMD3 > gpurun_out/p_plain3.log 2>&1 || { echo "plain C3 failed"; tail -3 gpurun_out/p_plain3.log; exit 1; }
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/r02_launches_C3.csv $CMD3 > gpurun_out/p_ncu1.log 2>&1; echo "launch list rc=$?"
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,sm__cycles_elapsed.max \
  --clock-control none -k regex:'k_fcol|k_srow' -s 6 -c 2 --csv --log-file gpurun_out/r02_l2_counters_C3_final.csv $CMD3 > gpurun_out/p_ncu2.log 2>&1; echo "counters rc=$?"
CMD2="python bench.py --workload C2 --algo fast --steps 1 --warmup 3 --hxv-only"
$CMD2 > gpurun_out/p_plain2.log 2>&1 &&
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_fcol|k_srow' -s 6 -c 2 -o gpurun_out/r02_full_C2 -f $CMD2 > gpurun_out/p_ncu3.log 2>&1; echo "full rc=$?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null | tail -2
