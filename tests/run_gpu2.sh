#!/bin/bash
# 2-GPU box: sharded parity under torchrun + the strong-scaling bench line at N=2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpu2.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_multi.log; tail -15 gpurun_out/pytest_multi.log
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_C3_n$N.json 2> gpurun_out/bench_C3_n$N.err; cat gpurun_out/bench_C3_n$N.json; tail -5 gpurun_out/bench_C3_n$N.err
