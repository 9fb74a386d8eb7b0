#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "hybrid or phonon or orbs or basis or hxv_matches" -p no:cacheprovider --timeout 600 --timeout-method=thread 2>&1 | tail -40 > gpurun_out/r02n_pytest.log; tail -25 gpurun_out/r02n_pytest.log
