#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "lanczos or chi or sector_scan or observ or gf_chains" -p no:cacheprovider --timeout 600 --timeout-method=thread 2>&1 | tail -30 > gpurun_out/r02k_pytest.log; tail -12 gpurun_out/r02k_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-c5 --no-fitted > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02k_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02k_bench.json").read().strip().splitlines()[-1])
print("hxv/s %.1f  lanczos it/s %.1f  chain %.1f" % (d["value"], d["lanczos_iter_per_s"], d["e2e"]["chain_hxv_per_s"]), d["e2e"].get("eigh"))
PY
