#!/bin/bash
# 2-GPU box: push-halo development run -- one-device emulation tests, the NCCL/IPC worker, bench lines with and without overlap
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "sharded" -p no:cacheprovider --timeout 500 --timeout-method=thread 2>&1 | tail -30 > gpurun_out/r02h_pytest.log; tail -8 gpurun_out/r02h_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tests/multigpu_worker.py > gpurun_out/r02h_multi$N.log 2>&1; echo "worker$N rc=$?"; grep -E "FAIL|MULTIGPU|rror" gpurun_out/r02h_multi$N.log | head -12
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/r02h_bench_n${N}$tag.json 2> gpurun_out/r02h_bench_n${N}$tag.err; echo "bench N=$N $tag rc=$?"; cut -c1-160 gpurun_out/r02h_bench_n${N}$tag.json; grep -E "rror|FAIL" gpurun_out/r02h_bench_n${N}$tag.err | head -3; }
run "" --no-c5 --no-fitted
run _w4 --no-c5 --no-fitted --opt halo_windows=4
run _noovl --no-c5 --no-fitted --opt no_overlap=1
run _h16 --no-c5 --no-fitted --opt halo_ctas=16
