#!/bin/bash
# 8-GPU box: the real NCCL + CUDA-IPC path at N ranks.  bench.py carries its own parity check (oracle columns on every
# rank); variants: one column window, no overlap (per-kernel split incl. the halo push).  FULL=1 adds the sharded
# parity worker (operator, Lanczos, GF chains incl. the spin-down exchange).
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/scale_gpus.txt 2>&1
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/scale_bench_n${N}$tag.json 2> gpurun_out/scale_bench_n${N}$tag.err; echo "bench N=$N $tag rc=$?"; cut -c1-160 gpurun_out/scale_bench_n${N}$tag.json; grep -E "rror|FAIL" gpurun_out/scale_bench_n${N}$tag.err | head -3; }
run "" --no-fitted
run _w1 --no-c5 --no-fitted --opt halo_windows=1
run _noovl --no-c5 --no-fitted --opt no_overlap=1
if [ -n "$FULL" ]; then
EDGPU_WORKER_QUICK=1 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29931 tests/multigpu_worker.py > gpurun_out/scale_multi$N.log 2>&1; echo "worker$N rc=$?"; grep -E "FAIL|MULTIGPU" gpurun_out/scale_multi$N.log | head -12
fi
