#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "phonon or orbs" -p no:cacheprovider --timeout 600 --timeout-method=thread 2>&1 | tail -40 > gpurun_out/r02m_pytest.log; tail -25 gpurun_out/r02m_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tests/multigpu_worker.py > gpurun_out/r02m_multi$N.log 2>&1; echo "worker$N rc=$?"; grep -E "FAIL|MULTIGPU|rror" gpurun_out/r02m_multi$N.log | head -12
