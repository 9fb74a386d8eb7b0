#!/bin/bash
# 8-GPU box: the real NCCL + CUDA-IPC path at 8 and 4 ranks (bench.py carries its own parity check), then the sharded
# parity worker (operator, Lanczos, GF chains incl. spin-down exchange) at 8 ranks.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/scale_gpus.txt 2>&1
for n in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_bench_n$n.json 2> gpurun_out/scale_bench_n$n.err; echo "bench N=$n rc=$?"; cut -c1-200 gpurun_out/scale_bench_n$n.json; tail -3 gpurun_out/scale_bench_n$n.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 8 --steps 10 --warmup 3 --no-c5 --no-fitted --opt no_overlap=1 > gpurun_out/scale_bench_n8_noovl.json 2> gpurun_out/scale_bench_n8_noovl.err; echo "bench N=8 no-overlap rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tests/multigpu_worker.py > gpurun_out/scale_multi8.log 2>&1; echo "worker8 rc=$?"; grep -E "FAIL|MULTIGPU" gpurun_out/scale_multi8.log | head -12
