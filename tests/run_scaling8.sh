#!/bin/bash
# 8-GPU box: the real NCCL + CUDA-IPC path.  bench.py carries its own parity check (oracle columns on every rank) and
# the serialised per-kernel split; then the NVLink microbenchmark of the halo traffic pattern, the sharded parity worker
# (operator, Lanczos, GF / susceptibility chains, phonons), the N = 4 and N = 2 lines on the same box, two variants.
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/scale_gpus.txt 2>&1
run() { n=$1; tag=$2; shift; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n --steps 10 --warmup 3 "$@" > gpurun_out/scale_bench_n${n}$tag.json 2> gpurun_out/scale_bench_n${n}$tag.err; echo "bench N=$n $tag rc=$?"; cut -c1-160 gpurun_out/scale_bench_n${n}$tag.json; grep -E "rror|FAIL" gpurun_out/scale_bench_n${n}$tag.err | head -3; }
run $N "" --no-fitted
timeout 120 tools/nvlink_push.bin $N 12870 2570 > gpurun_out/nvlink_push_n$N.log 2>&1; echo "nvlink microbench rc=$?"; grep -E "ctas  148|ctas   32" gpurun_out/nvlink_push_n$N.log | head -10
EDGPU_WORKER_QUICK=1 timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29931 tests/multigpu_worker.py > gpurun_out/scale_multi$N.log 2>&1; echo "worker$N rc=$?"; grep -E "FAIL|MULTIGPU" gpurun_out/scale_multi$N.log | head -12
if [ "$N" = "8" ]; then
run 4 "" --no-c5 --no-fitted
run 2 "" --no-c5 --no-fitted
run 8 _h64 --no-c5 --no-fitted --opt halo_ctas=64
run 8 _w8 --no-c5 --no-fitted --opt halo_windows=8
fi
