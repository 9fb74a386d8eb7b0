import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dmft-lanc-ed_b200"))
import numpy as np
import edgpu
from edgpu import configs
name, nup, ndw = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
cfg = configs.config(name)
s = edgpu.Solver(ed_sparse_h=(sys.argv[4] == "1"), device=0, **configs.solver_kwargs(cfg))
s.set_option("hxv_algo", edgpu.ALGO_FAST)
for kv in sys.argv[5:]:
    k, v = kv.split("="); s.set_option(k, int(v))
s.build_Hv_sector(s.get_sector(nup, ndw))
v = configs.bench_vector(s.nloc); v /= np.linalg.norm(v)
hv = s.spHtimesV(v)
print("ok", np.abs(hv).max())
