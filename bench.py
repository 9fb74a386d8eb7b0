#!/usr/bin/env python
"""bench.py -- Lanczos H*v throughput of the B200 engine (and of the CPU reference arm).

    python bench.py --gpus N --steps K --warmup W [--workload C3] [--stored] [--algo auto|gather|tiled]
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one H*v of the named sector on a synthetic vector v_i = sin(0.37 i) + 0.1
(SURVEY.md 8d).  Default workload: BASELINE.json configs[2] (single-band Hubbard, Ns=16, sector
8:8, dim 165 636 900, 1.3 GB/vector), the configuration the metric is quoted on at 1/2/4/8 B200;
--workload C5 selects the Ns=18 9:9 sector.  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "dmft-lanc-ed_b200"))
# stdout carries exactly one JSON line: NCCL's own banner ("NCCL version ...") goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

METRIC = "lanczos_hxv_per_s"
UNIT = "H*v/s"
WORKLOADS = {
    "C1": "single-band Hubbard Norb=1 Nbath=7 (Ns=8), sector 4:4, dim 4900",
    "C2": "single-band Hubbard Norb=1 Nbath=13 (Ns=14), sector 7:7, dim 11778624",
    "C3": "single-band Hubbard Norb=1 Nbath=15 (Ns=16), sector 8:8, dim 165636900",
    "C4": "two-orbital Hubbard-Kanamori Norb=2 Nbath=4 (Ns=10), sector 5:5, dim 63504",
    "C5": "single-band Hubbard Norb=1 Nbath=17 (Ns=18), sector 9:9, dim 2363904400",
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=os.environ.get("EDGPU_BENCH_WORKLOAD", "C3"))
    ap.add_argument("--stored", action="store_true", help="ED_SPARSE_H=T: stream spH0d (24 B/element) instead of recomputing the diagonal (16 B/element)")
    ap.add_argument("--algo", default="auto", choices=["auto", "gather", "tiled", "fast"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--hxv-only", action="store_true", help="only the device-resident H*v loop (used under ncu)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--opt", action="append", default=[], help="engine option key=value (edgpu_set_option), repeatable")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if "Active" in val and "Not" not in val:
                    reasons.add(name)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


# ------------------------------------------------------------------------------------------------
# CPU arm: bounded sample of the restated stored-sparse path on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_sample(workload, seconds, max_threads=None):
    """Times the oracle's spMatVec_main loop nests (the restated reference CPU path, kind="port":
    the Fortran reference cannot be compiled here) on column blocks of the workload's sector, one
    block per thread = the per-rank work of the reference's MPI split, communication excluded."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    from edgpu import configs

    cfg = configs.config(workload)
    nup, ndw = cfg["nup"], cfg["ndw"]
    o = O.Oracle(**configs.solver_kwargs(cfg))
    cores = os.cpu_count() or 1
    threads = min(cores, max_threads) if max_threads else cores
    ns = o.ns
    import math
    dimup, dimdw = math.comb(ns, nup), math.comb(ns, ndw)
    dim = dimup * dimdw
    nblk = max(threads, min(dimdw, max(threads, int(dim / 2.0e6))))     # ~2e6 elements per block
    v = configs.bench_vector(dim)
    secs = [o.sector(nup, ndw, r, nblk) for r in range(threads)]
    outs = [np.empty(s.nloc) for s in secs]
    t0 = time.perf_counter()
    n1 = O.spmatvec_blocks_mt(secs, v, outs, threads)
    t1 = time.perf_counter() - t0
    passes = int(max(1, min(200, seconds / max(t1, 1e-6))))
    t0 = time.perf_counter()
    nel = 0
    for _ in range(passes):
        nel += O.spmatvec_blocks_mt(secs, v, outs, threads)
    dt = time.perf_counter() - t0
    rate = nel / dt
    for s in secs:
        s.close()
    return {"value": rate / dim, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d passes over %d of %d i_dw column blocks (%d elements each pass) of %s, spMatVec_main loop "
                      "nests restated in C (gcc -O3), one block per thread, no inter-rank transpose; %.1f s"
                      % (passes, threads, nblk, n1, workload, dt),
            "elements_per_s": rate, "seconds": dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    per_step = max(0.5, min(10.0, 120.0 / (steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_sample(args.workload, 0.2)
    vals = []
    t0 = time.perf_counter()
    last = None
    for _ in range(steps):
        last = cpu_sample(args.workload, per_step)
        vals.append(last["value"])
    dt = time.perf_counter() - t0
    value = float(np.mean(vals))
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 / value, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOADS[args.workload], "operator": "stored sparse (ED_SPARSE_H=T), CPU"},
            "cpu_baseline": dict(last, value=value),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": dt}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import edgpu
    from edgpu import configs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if edgpu.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch = None
    if world > 1 or not args.hxv_only:                           # torch = distributed plumbing + pinned host buffers only
        import torch
        torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        if torch is not None:
            torch.cuda.synchronize()

    cfg = configs.config(args.workload)
    nup, ndw = cfg["nup"], cfg["ndw"]
    s = edgpu.Solver(ed_sparse_h=args.stored, device=local, **configs.solver_kwargs(cfg))
    s.set_option("hxv_algo", {"auto": edgpu.ALGO_AUTO, "gather": edgpu.ALGO_GATHER, "tiled": edgpu.ALGO_TILED, "fast": edgpu.ALGO_FAST}[args.algo])
    for kv in args.opt:
        k, v = kv.split("=")
        s.set_option(k, int(v))
    if world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(edgpu.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        s.set_comm(rank, world, bytes(idt.cpu().numpy().tobytes()))
    isec = s.get_sector(nup, ndw)
    s.build_Hv_sector(isec)
    nloc, dim = s.nloc, s.dimup * s.dimdw
    d_v, d_hv = s.dev_alloc(8 * nloc), s.dev_alloc(8 * nloc)
    s.dev_fill_bench_vector(d_v, nloc, s.ishift)
    s.sync()

    # ---- device-resident H*v: W warm-ups, then exactly K timed steps -------------------------------
    for _ in range(max(3, args.warmup)):
        s.hxv_device(d_v, d_hv)
    s.sync()
    barrier()
    clocks = Clocks(local) if rank == 0 else None
    l0 = s.launch_count()
    barrier()
    ms = s.time_hxv_device(d_v, d_hv, args.steps)              # CUDA events on the engine's stream
    s.sync()
    barrier()
    launches = s.launch_count() - l0
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_step = ms / args.steps
    value = 1000.0 / ms_step
    if args.hxv_only:
        kp = s.time_hxv_passes(d_v, d_hv, max(3, min(args.steps, 10))) if world == 1 else []
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "ms_per_step": ms_step, "hxv_only": True,
                              "gpu_launches": int(launches), "kernels": kp}))
        s.delete_Hv_sector()
        s.close()
        return

    # ---- per-kernel split of the same loop (single rank; separate, untimed-for-value pass) ---------
    passes = []
    if world == 1:
        try:
            passes = s.time_hxv_passes(d_v, d_hv, max(3, min(args.steps, 10)))
        except Exception:
            passes = []

    # ---- Lanczos iterations per second (device-resident recurrence, no host sync inside) ------------
    s.time_lanczos_device(d_v, 3)
    barrier()
    ms_l = s.time_lanczos_device(d_v, args.steps)
    barrier()
    if dist is not None:
        t = torch.tensor([ms_l], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_l = float(t.item())
    clk = clocks.stop() if clocks else None

    # ---- end to end through the reference-facing call spHtimesV_p(Nloc, v, Hv): host buffers -------
    hv_h = torch.empty(nloc, dtype=torch.float64).pin_memory()
    v_h = torch.empty(nloc, dtype=torch.float64).pin_memory()
    v_h.numpy()[:] = configs.bench_vector(nloc, s.ishift)
    ke = max(1, min(args.steps, 8))
    s.spHtimesV_ptr(nloc, v_h.data_ptr(), hv_h.data_ptr())
    barrier()
    t0 = time.perf_counter()
    for _ in range(ke):
        s.spHtimesV_ptr(nloc, v_h.data_ptr(), hv_h.data_ptr())  # H2D + H*v + D2H, synchronous
    barrier()
    e2e_s = (time.perf_counter() - t0) / ke
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    # e2e of a whole chain: one sp_lanc_tridiag call (H2D of the start vector once, coefficients back)
    nl = max(2, min(args.steps, 50))
    barrier()
    t0 = time.perf_counter()
    s.sp_lanc_tridiag(v_h.numpy(), nl)
    barrier()
    chain_s = time.perf_counter() - t0

    bytes_per_el = 24 if args.stored else 16
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    step_gbs = bytes_per_el * nloc / (ms_step * 1e-3) / 1e9      # per GPU: algorithmic bytes of the local shard
    # per-kernel split (CUDA events on the engine's stream between the launches) and the dominant kernel
    kernels = []
    if passes:
        # minimum bytes each launch has to move per element: row kernel reads x and writes y (+ streamed spH0d),
        # column kernel reads x, reads y, writes y; the whole H*v step is judged against 16 (24) B/element
        alg = {"k_srow": 16 + (8 if args.stored else 0), "k_fcol": 24, "k_tile_col": 16 + (8 if args.stored else 0),
               "k_tile_row": 24, "k_hxv_gather": bytes_per_el, "k_halo_pull": 0}
        tot = sum(ms_k for _, ms_k in passes)
        for name, ms_k in passes:
            b = alg.get(name, bytes_per_el)
            kernels.append({"name": name, "ms": ms_k, "share": ms_k / tot, "algorithmic_bytes_per_element": b,
                            "achieved_gbs": b * nloc / (ms_k * 1e-3) / 1e9, "frac": b * nloc / (ms_k * 1e-3) / 1e9 / peak})
    traffic = None
    dom = max(kernels, key=lambda k: k["ms"]) if kernels else None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_dram_traffic.json")))
        key = "%s_%s" % (args.workload, "stored" if args.stored else "direct")
        if dom and world == 1 and key in tr and dom["name"] in tr[key]:
            traffic = tr[key][dom["name"]]["dram_bytes_per_launch"]
    except Exception:
        pass
    if dom:
        roof = {"bound": "hbm", "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": dom["frac"], "traffic": traffic,
                "kernel": dom["name"], "ms_per_launch": dom["ms"],
                "algorithmic_bytes_per_launch": dom["algorithmic_bytes_per_element"] * nloc,
                "algorithmic_bytes_per_element": dom["algorithmic_bytes_per_element"]}
    else:
        roof = {"bound": "hbm", "achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak, "traffic": None,
                "kernel": "whole H*v step (all kernels of one step), per GPU",
                "algorithmic_bytes_per_element": bytes_per_el}
    roof.update({"peak_source": peak_src, "elements_per_launch": nloc,
                 "step": {"algorithmic_bytes_per_element": bytes_per_el, "achieved": step_gbs, "frac": step_gbs / peak,
                          "note": "north-star roofline: read v + write Hv%s per element over the whole H*v (all launches)"
                                  % (" + stream spH0d" if args.stored else "; diagonal recomputed")},
                 "kernels": kernels})
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "operator": "stored (ED_SPARSE_H=T)" if args.stored else "direct (ED_SPARSE_H=F, diagonal recomputed)",
                   "algo": args.algo, "l2": "input vector %.0f MB per GPU, larger than the 126 MB L2; no flush" % (8 * nloc / 1e6),
                   "sharding": "i_dw columns, %d rank(s)" % world, "vector": "v_i = sin(0.37 i) + 0.1"},
        "lanczos_iter_per_s": 1000.0 * args.steps / ms_l,
        "roofline": roof,
        "e2e": {"value": 1.0 / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 8 * nloc * world, "d2h_bytes_per_step": 8 * nloc * world,
                "call": "edgpu_hxv == spHtimesV_p(Nloc,v,Hv) with pinned host arrays",
                "chain_hxv_per_s": nl / chain_s, "chain_call": "edgpu_sp_lanc_tridiag, %d steps, start vector from host" % nl},
        "gpu_launches": int(launches),
        "clocks": clk,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_sample(args.workload, args.cpu_seconds)
        except Exception as e:                                   # the baseline never blocks the GPU line
            line["cpu_baseline"] = {"error": repr(e)}
    s.dev_free(d_v)
    s.dev_free(d_hv)
    s.delete_Hv_sector()
    s.close()
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    # stdout carries exactly ONE JSON line: everything else any library writes to fd 1 (NCCL's version banner,
    # warnings) is sent to stderr for the whole run, and the line goes to a private copy of the real stdout
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    _print = print

    def print(*args, **kw):                                      # noqa: A001 -- the module's only stdout writer
        kw.setdefault("file", _OUT)
        _print(*args, **kw)
        _OUT.flush()

    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
