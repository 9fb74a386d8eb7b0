#!/bin/bash
# 1-GPU box: fused Lanczos step -- Lanczos / GF / sector-scan parity tests, bench lines with and without it
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "lanczos or tridiag or gf or sector_scan or observ or krylov" -p no:cacheprovider --timeout 600 --timeout-method=thread 2>&1 | tail -30 > gpurun_out/r02i_pytest.log; tail -12 gpurun_out/r02i_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-c5 > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02i_bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-c5 --no-fitted --opt no_lz=1 > gpurun_out/r02i_bench_nolz.json 2> gpurun_out/r02i_bench_nolz.err; echo "bench no_lz rc=$?"
python - <<'PY'
import json
for f in ["r02i_bench.json","r02i_bench_nolz.json"]:
    try:
        d=json.loads(open("gpurun_out/"+f).read().strip().splitlines()[-1])
        print(f, "hxv/s %.1f  lanczos it/s %.1f  chain %.1f" % (d["value"], d["lanczos_iter_per_s"], d["e2e"]["chain_hxv_per_s"]), [(k["name"], round(k["ms"],3)) for k in d["roofline"]["kernels"]])
        if d.get("fitted_bath"): print("  fitted", d["fitted_bath"]["hxv_per_s"], d["fitted_bath"]["lanczos_iter_per_s"])
    except Exception as e: print(f, "ERR", e)
PY
