#!/bin/bash
# GPU call B of round 2 (2 GPUs): real NCCL + IPC path (halo pull over NVLink, lists, sharded GF incl. spin-down), bench at N=2 and N=1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/r02b_gpu.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_worker.py > gpurun_out/r02b_multi.log 2>&1; echo "multi rc=$?"; grep -E "FAIL|MULTIGPU" gpurun_out/r02b_multi.log | head -20; tail -3 gpurun_out/r02b_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02b_bench_n2.json 2> gpurun_out/r02b_bench_n2.err; echo "bench2 rc=$?"; cut -c1-1500 gpurun_out/r02b_bench_n2.json; tail -5 gpurun_out/r02b_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --no-c5 --no-fitted --opt no_overlap=1 > gpurun_out/r02b_bench_n2_noovl.json 2> gpurun_out/r02b_bench_n2_noovl.err; echo "bench2 no-overlap rc=$?"; cut -c1-400 gpurun_out/r02b_bench_n2_noovl.json
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; echo "bench1 rc=$?"; cut -c1-3000 gpurun_out/r02b_bench_n1.json; tail -5 gpurun_out/r02b_bench_n1.err
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "spin_down or c5_ns18 or lanczos or full_size_properties" -p no:cacheprovider --timeout 500 --timeout-method=thread 2>&1 | tail -30 > gpurun_out/r02b_pytest_new.log; tail -6 gpurun_out/r02b_pytest_new.log
