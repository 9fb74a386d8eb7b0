#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "sector_scan or lanczos_ground or observables or gf_chains" -p no:cacheprovider --timeout 500 --timeout-method=thread 2>&1 | tail -30 > gpurun_out/r02g_pytest.log; tail -8 gpurun_out/r02g_pytest.log
