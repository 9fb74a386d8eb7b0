"""bench.py's JSON line, assembled on the CPU from canned measurements (the device work is stubbed out): the keys of the
driver contract are present and consistent, both arms carry the same `config.operator`, the dominant-kernel roofline is
the slower of the two passes, and the halo block of a multi-rank line survives.  Guards the reporting code, which only
runs on a GPU box otherwise."""
import argparse
import io
import json
import os
import sys
from contextlib import redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _args(**over):
    a = dict(gpus=1, steps=20, warmup=3, workload="C3", stored=False, algo="auto", hxv_only=False, no_fitted=True, no_c5=True,
             no_cpu_baseline=True, cpu_seconds=1.0, opt=[], impl="b200")
    a.update(over)
    return argparse.Namespace(**a)


def _canned(world):
    nloc = 165636900 // world
    res = {"workload": "C3", "ms_per_step": 2.0, "value": 500.0, "launches": 40, "nloc": nloc, "dim": 165636900, "build_s": 0.1,
           "passes": [("k_srow", 1.06), ("k_fcol", 0.94)], "lanczos_ms_per_iter": 3.0, "e2e_s": 0.05, "chain_s": 0.08, "chain_steps": 20,
           "parity_check": {"ok": True, "tolerance": 1e-12}, "eigh": {"seconds": 0.58, "nlanc": 79, "hxv": 158, "e0": -17.3, "call": "x"}}
    if world > 1:
        res["halo"] = {"bytes_out_per_gpu": 3.47e8, "bytes_in_per_gpu": 3.47e8, "windows": 4, "push_ms": 0.7,
                       "serialised_passes": [("k_halo_push", 0.7), ("k_srow", 0.12), ("k_halo_wait", 0.09), ("k_fcol", 0.29)],
                       "nvlink_out_gbs_per_gpu": 500.0}
    return res


def _run(monkeypatch, world):
    import bench

    class H:
        def __init__(self, args):
            self.args, self.world, self.rank, self.local, self.numa, self.dist, self.torch = args, world, 0, 0, None, None, None

        def init(self):
            pass

        def barrier(self):
            pass

    class Clk:
        def __init__(self, local):
            pass

        def stop(self):
            return {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": [], "samples": 10}

    monkeypatch.setattr(bench, "Harness", H)
    monkeypatch.setattr(bench, "Clocks", Clk)
    monkeypatch.setattr(bench, "measure", lambda Hh, wl, steps, full: _canned(world))
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.run_b200(_args(gpus=world))
    lines = [l for l in buf.getvalue().splitlines() if l.strip()]
    assert len(lines) == 1                                          # exactly one JSON line on stdout
    return json.loads(lines[0])


def test_line_of_the_gpu_arm(monkeypatch):
    import bench
    d = _run(monkeypatch, 1)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "roofline", "e2e", "gpu_launches", "clocks", "parity_check"):
        assert k in d, k
    assert d["metric"] == "lanczos_hxv_per_s" and d["unit"] == "H*v/s" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["scaling"] == "strong" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["config"]["operator"] == bench.OPERATOR and "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["kernel"] == "k_srow" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert abs(r["achieved"] - 16 * 165636900 / 1.06e-3 / 1e9) < 1e-6 * r["achieved"]
    assert {k["name"] for k in r["kernels"]} == {"k_srow", "k_fcol"}
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == e["d2h_bytes_per_step"] == 8 * 165636900 and abs(e["value"] - 20.0) < 1e-9
    assert e["eigh"]["nlanc"] == 79 and abs(e["chain_hxv_per_s"] - 250.0) < 1e-9
    assert "halo" not in d


def test_line_of_a_multi_rank_run_keeps_the_halo_block(monkeypatch):
    d = _run(monkeypatch, 8)
    assert d["n_gpus"] == 8 and d["halo"]["windows"] == 4 and d["halo"]["nvlink_out_gbs_per_gpu"] == 500.0
    assert d["roofline"]["traffic"] is None                        # the ncu DRAM figure is a single-GPU capture
    assert d["config"]["sharding"].startswith("i_dw columns, 8")
