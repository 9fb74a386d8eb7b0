#!/bin/bash
# 8-GPU box: the real NCCL + CUDA-IPC path at 8 ranks (bench.py carries its own parity check), then a short subset of
# the sharded parity worker (operator, Lanczos, GF chains incl. the spin-down exchange).
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/scale_gpus.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_bench_n$N.json 2> gpurun_out/scale_bench_n$N.err; echo "bench N=$N rc=$?"; cut -c1-200 gpurun_out/scale_bench_n$N.json; tail -3 gpurun_out/scale_bench_n$N.err
EDGPU_WORKER_QUICK=1 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tests/multigpu_worker.py > gpurun_out/scale_multi$N.log 2>&1; echo "worker$N rc=$?"; grep -E "FAIL|MULTIGPU" gpurun_out/scale_multi$N.log | head -12
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus $N --steps 10 --warmup 3 --no-c5 --no-fitted --opt no_overlap=1 > gpurun_out/scale_bench_n${N}_noovl.json 2> gpurun_out/scale_bench_n${N}_noovl.err; echo "bench N=$N no-overlap rc=$?"
