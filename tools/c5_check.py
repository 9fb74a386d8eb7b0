"""Full-size Ns=18 (C5, dim 2 363 904 400, 19 GB per vector) on ONE B200: fast path (k_srow + 2-CTA cluster column
kernel) against the one-pass gather kernel on the whole vector, then timings."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dmft-lanc-ed_b200"))
import numpy as np
import edgpu
from edgpu import configs
cfg = configs.config("C5")
s = edgpu.Solver(ed_sparse_h=False, device=0, **configs.solver_kwargs(cfg))
t0 = time.time()
s.build_Hv_sector(s.get_sector(cfg["nup"], cfg["ndw"]))
n = s.nloc
print("build %.1f s, nloc %d" % (time.time() - t0, n), flush=True)
dx, dy = s.dev_alloc(8 * n), s.dev_alloc(8 * n)
s.dev_fill_bench_vector(dx, n, 0)
out = {}
s.set_option("hxv_algo", edgpu.ALGO_FAST)
s.hxv_device(dx, dy); s.sync()
res = {"fast_ms": s.time_hxv_device(dx, dy, 3) / 3, "fast_kernels": s.time_hxv_passes(dx, dy, 2)}
yf = np.empty(n); s.dev_download(dy, yf)
s.set_option("hxv_algo", edgpu.ALGO_GATHER)
s.hxv_device(dx, dy); s.sync()
res["gather_ms"] = s.time_hxv_device(dx, dy, 2) / 2
yg = np.empty(n); s.dev_download(dy, yg)
scale = float(np.abs(yg[::97]).max())
err = 0.0
for a in range(0, n, 1 << 27):
    err = max(err, float(np.abs(yf[a:a + (1 << 27)] - yg[a:a + (1 << 27)]).max()))
res.update({"max_abs_diff": err, "scale": scale, "rel": err / scale, "hxv_per_s_fast": 1000.0 / res["fast_ms"],
            "roofline_frac_16B": 16.0 * n / (res["fast_ms"] * 1e-3) / 1e9 / 6546.6})
print(json.dumps(res), flush=True)
assert err < 1e-12 * scale
s.set_option("hxv_algo", edgpu.ALGO_AUTO)
ms_l = s.time_lanczos_device(dx, 3)
print(json.dumps({"lanczos_iter_per_s": 3000.0 / ms_l}), flush=True)
