#!/usr/bin/env python
"""bench.py -- Lanczos H*v throughput of the B200 engine (and of the CPU reference arm).

    python bench.py --gpus N --steps K --warmup W [--workload C3] [--stored] [--algo auto|gather|tiled|fast]
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one H*v of the named sector on a synthetic vector v_i = sin(0.37 i) + 0.1
(SURVEY.md 8d).  Default workload: BASELINE.json configs[2] (single-band Hubbard, Ns=16, sector
8:8, dim 165 636 900, 1.3 GB/vector), the configuration the metric is quoted on at 1/2/4/8 B200.
The same line carries "parity_check" (the result of the timed operator against the CPU oracle on
columns at both ends and the middle of every rank's shard, plus global inner products that must
not depend on the rank count) -- the run exits non-zero when it fails -- and "c5": the Ns=18 9:9
sector (dim 2 363 904 400, 19 GB per vector) sharded over the same GPUs.  One JSON line is printed
by rank 0.
"""
import argparse
import glob
import json
import math
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
OPERATOR = "spHtimesV_p: Hv = Hd*v + (Hdw x 1)v + (1 x Hup)v on the whole sector"   # config.operator of BOTH arms
WORKLOADS = {
    "C1": "single-band Hubbard Norb=1 Nbath=7 (Ns=8), sector 4:4, dim 4900",
    "C2": "single-band Hubbard Norb=1 Nbath=13 (Ns=14), sector 7:7, dim 11778624",
    "C3": "single-band Hubbard Norb=1 Nbath=15 (Ns=16), sector 8:8, dim 165636900",
    "C3V": "single-band Hubbard Norb=1 Nbath=15 (Ns=16), sector 8:8, dim 165636900, fitted bath (level-dependent V_k)",
    "C4": "two-orbital Hubbard-Kanamori Norb=2 Nbath=4 (Ns=10), sector 5:5, dim 63504",
    "C5": "single-band Hubbard Norb=1 Nbath=17 (Ns=18), sector 9:9, dim 2363904400",
}
PARITY_TOL = 1e-12           # bench gate (the -m gpu tests use 1e-13): |y - y_oracle| / max|y_oracle| per checked column


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
    ap.add_argument("--no-c5", action="store_true", help="skip the Ns=18 sub-measurement")
    ap.add_argument("--no-fitted", action="store_true", help="skip the fitted-bath (level-dependent V_k) sub-measurement")
    ap.add_argument("--hxv-only", action="store_true", help="only the device-resident H*v loop (used under ncu)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--opt", action="append", default=[], help="engine option key=value (edgpu_set_option), repeatable")
    return ap.parse_args()


def workload_config(name):
    from edgpu import configs
    if name == "C3V":
        return configs.config("NS16V")
    return configs.config(name)


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
# CPU arm: the restated reference path on the host cores
# ------------------------------------------------------------------------------------------------
class CpuReference:
    """spMatVec_MPI_main (ED_HAMILTONIAN_SPARSE_HxV.f90:568-694) of the workload's sector as the oracle restates
    it (kind="port": the Fortran reference cannot be compiled here): the reference's own i_dw split over P ranks,
    one rank per host thread, INCLUDING the two vector_transpose_MPI exchanges, on the whole vector."""

    def __init__(self, workload, threads=None):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle as O
        from edgpu import configs
        self.O = O
        cfg = workload_config(workload)
        self.nup, self.ndw = cfg["nup"], cfg["ndw"]
        self.o = O.Oracle(**configs.solver_kwargs(cfg))
        cores = os.cpu_count() or 1
        self.threads = min(cores, threads) if threads else cores
        ns = self.o.ns
        self.dimup, self.dimdw = math.comb(ns, self.nup), math.comb(ns, self.ndw)
        self.dim = self.dimup * self.dimdw
        self.P = max(1, min(self.threads, self.dimdw))
        t0 = time.perf_counter()
        self.secs = [self.o.sector(self.nup, self.ndw, r, self.P) for r in range(self.P)]
        self.build_s = time.perf_counter() - t0
        self.v = configs.bench_vector(self.dim)
        self.hv = np.empty(self.dim)

    def step(self):
        t0 = time.perf_counter()
        self.O.spmatvec_mpi_prebuilt(self.secs, self.v, self.hv, self.threads)
        return time.perf_counter() - t0

    def single_thread(self, frac_blocks=8):
        """spMatVec_main (the serial routine, :391-485) on ONE thread, timed on one of `frac_blocks` column blocks."""
        with self.o.sector(self.nup, self.ndw, frac_blocks // 2, frac_blocks) as blk:
            out = [np.empty(blk.nloc)]
            t0 = time.perf_counter()
            n = self.O.spmatvec_blocks_mt([blk], self.v, out, 1)
            dt = time.perf_counter() - t0
        return {"value": n / dt / self.dim, "unit": UNIT, "cores": 1,
                "sample": "one of %d i_dw column blocks (%d elements), serial spMatVec_main loop nests, %.2f s" % (frac_blocks, n, dt)}

    def close(self):
        for s in self.secs:
            s.close()
        self.secs = []


def cpu_baseline(workload, seconds, steps_max=50):
    ref = CpuReference(workload)
    ref.step()
    times = []
    t0 = time.perf_counter()
    while len(times) < steps_max and (time.perf_counter() - t0 < seconds or len(times) < 2):
        times.append(ref.step())
    one = ref.single_thread()
    dt = float(np.sum(times))
    out = {"value": len(times) / dt, "unit": UNIT, "cores": ref.threads, "kind": "port",
           "sample": "%d full H*v of %s through the restated spMatVec_MPI_main (gcc -O3): %d ranks on %d host threads, whole vector, "
                     "both vector_transpose_MPI exchanges included; %.1f s (+ %.1f s build_Hv_sector, not timed)"
                     % (len(times), workload, ref.P, ref.threads, dt, ref.build_s),
           "seconds": dt, "single_thread": one}
    ref.close()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref = CpuReference(args.workload)
    for _ in range(max(1, args.warmup)):
        ref.step()
    steps = max(1, args.steps)
    times = [ref.step() for _ in range(steps)]
    dt = float(np.sum(times))
    value = steps / dt
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": max(1, args.warmup),
            "ms_per_step": 1000.0 * dt / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOADS[args.workload], "operator": OPERATOR, "vector": "v_i = sin(0.37 i) + 0.1"},
            "operator_form": "stored sparse (ED_SPARSE_H=T), CPU",
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": ref.threads, "kind": "port",
                             "sample": "every step = one full H*v of %s through the restated spMatVec_MPI_main (gcc -O3): %d ranks on %d host "
                                       "threads, both vector_transpose_MPI exchanges included" % (args.workload, ref.P, ref.threads)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": dt}
    ref.close()
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def pin_to_gpu_numa_node(local):
    """Best effort: run this process (and first-touch its pinned staging buffers) on the NUMA node the GPU hangs off."""
    try:
        bus = subprocess.check_output(["nvidia-smi", "-i", str(local), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                      stderr=subprocess.DEVNULL, timeout=20).decode().strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


class Harness:
    def __init__(self, args):
        self.args = args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.torch = None
        self.dist = None

    def init(self):
        import edgpu
        if edgpu.device_count() < 1:
            raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
        self.numa = pin_to_gpu_numa_node(self.local)
        if self.world > 1 or not self.args.hxv_only:             # torch = distributed plumbing + pinned host buffers only
            import torch
            self.torch = torch
            torch.cuda.set_device(self.local)
        if self.world > 1:
            import torch.distributed as dist
            self.dist = dist
            dist.init_process_group("nccl", device_id=self.torch.device("cuda", self.local))

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        if self.torch is not None:
            self.torch.cuda.synchronize()

    def maxr(self, x):
        if self.dist is None:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def minr(self, x):
        return -self.maxr(-x)

    def uid(self):
        import edgpu
        idt = self.torch.zeros(128, dtype=self.torch.uint8, device="cuda")
        if self.rank == 0:
            idt.copy_(self.torch.frombuffer(bytearray(edgpu.comm_unique_id()), dtype=self.torch.uint8))
        self.dist.broadcast(idt, 0)
        return bytes(idt.cpu().numpy().tobytes())


def parity_check(H, s, cfg, d_v, d_hv, hermiticity):
    """d_hv = H d_v (just computed by the timed operator) against the CPU oracle on this rank's first, middle and last
    local i_dw column (spMatVec_main loop nests, oracle/), plus <y,y>, <x,y> all-reduced over the ranks (they must not
    depend on the rank count) and, when memory allows, <x,Hz> = <Hx,z>."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    from edgpu import configs
    out = {"tolerance": PARITY_TOL}
    out["yy"] = s.dev_dot(d_hv, d_hv)
    out["xy"] = s.dev_dot(d_v, d_hv)
    o = O.Oracle(**configs.solver_kwargs(cfg))
    c0 = s.ishift // s.dimup
    cols = sorted(set([c0, c0 + s.qdw // 2, c0 + s.qdw - 1]))
    worst = 0.0
    for j in cols:
        with o.sector(cfg["nup"], cfg["ndw"], j, s.dimdw) as blk:            # "rank j of DimDw ranks" = column j alone
            ci = blk.block_columns()
            xc = np.stack([configs.bench_vector(s.dimup, int(q) * s.dimup) for q in ci])
            ref = blk.spmatvec_block_cols(ci, xc)
        got = np.empty(s.dimup)
        s.dev_download_slice(d_hv, (j - c0) * s.dimup, got)
        worst = max(worst, float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300)))
    out["oracle_columns_per_rank"] = len(cols)
    out["oracle_max_rel_err"] = H.maxr(worst)
    ok = worst < PARITY_TOL
    if hermiticity:
        d_z, d_hz = s.dev_alloc(8 * s.nloc), s.dev_alloc(8 * s.nloc)
        s.dev_fill_bench_vector(d_z, s.nloc, s.ishift + 12345)
        s.hxv_device(d_z, d_hz)
        xhz, hxz = s.dev_dot(d_v, d_hz), s.dev_dot(d_hv, d_z)
        scale = math.sqrt(s.dev_dot(d_v, d_v) * s.dev_dot(d_hz, d_hz))
        out["hermiticity_rel"] = abs(xhz - hxz) / max(scale, 1e-300)
        ok = ok and out["hermiticity_rel"] < 1e-10
        s.dev_free(d_z)
        s.dev_free(d_hz)
    out["ok"] = bool(H.minr(1.0 if ok else 0.0) > 0.5)
    return out


def dram_traffic(workload, stored, kernel):
    """Per-launch DRAM bytes of `kernel` from this round's ncu capture (profiles/rNN_dram_traffic.json, newest round)."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_dram_traffic.json")))
    for f in reversed(files):
        try:
            tr = json.load(open(f))
            key = "%s_%s" % (workload, "stored" if stored else "direct")
            if key in tr and kernel in tr[key]:
                return tr[key][kernel]["dram_bytes_per_launch"], os.path.basename(f)
        except Exception:
            continue
    return None, None


def measure(H, workload, steps, full):
    """One sector: build, W warm-ups, `steps` timed device-resident H*v, per-kernel split, parity check; with `full`
    also the Lanczos loop and the end-to-end calls."""
    import edgpu
    from edgpu import configs
    args = H.args
    world, rank, local = H.world, H.rank, H.local
    cfg = workload_config(workload)
    nup, ndw = cfg["nup"], cfg["ndw"]
    s = edgpu.Solver(ed_sparse_h=args.stored, device=local, **configs.solver_kwargs(cfg))
    s.set_option("hxv_algo", {"auto": edgpu.ALGO_AUTO, "gather": edgpu.ALGO_GATHER, "tiled": edgpu.ALGO_TILED, "fast": edgpu.ALGO_FAST}[args.algo])
    for kv in args.opt:
        k, v = kv.split("=")
        s.set_option(k, int(v))
    if world > 1:
        s.set_comm(rank, world, H.uid())
    t0 = time.perf_counter()
    s.build_Hv_sector(s.get_sector(nup, ndw))
    nloc, dim = s.nloc, s.dimup * s.dimdw
    d_v, d_hv = s.dev_alloc(8 * nloc), s.dev_alloc(8 * nloc)
    s.dev_fill_bench_vector(d_v, nloc, s.ishift)
    s.hxv_device(d_v, d_hv)                                       # first application builds the kernel plan
    s.sync()
    build_s = time.perf_counter() - t0

    # ---- device-resident H*v: W warm-ups, then exactly K timed steps -------------------------------
    for _ in range(max(3, args.warmup)):
        s.hxv_device(d_v, d_hv)
    s.sync()
    H.barrier()
    l0 = s.launch_count()
    H.barrier()
    ms = s.time_hxv_device(d_v, d_hv, steps)                    # CUDA events on the engine's stream
    s.sync()
    H.barrier()
    launches = s.launch_count() - l0
    ms_step = H.maxr(ms) / steps
    res = {"workload": workload, "ms_per_step": ms_step, "value": 1000.0 / ms_step, "launches": int(launches),
           "nloc": nloc, "dim": dim, "build_s": build_s}
    if args.hxv_only:
        res["passes"] = s.time_hxv_passes(d_v, d_hv, max(3, min(steps, 10)))
        s.delete_Hv_sector()
        s.close()
        return res

    # ---- per-kernel split of the same loop (separate, untimed-for-value pass; rank 0's numbers) ------
    try:
        res["passes"] = s.time_hxv_passes(d_v, d_hv, max(3, min(steps, 10)))
    except Exception:
        res["passes"] = []
    H.barrier()
    if H.world > 1:
        # sharded: the halo push runs on its own stream next to the row pass, so the split above shows the waits, not
        # the kernels.  Run the same H*v serialised once more (no_overlap) for the true per-kernel durations and the
        # NVLink rate of the push (collective: every rank does it).
        try:
            bo, bi, nwin = s.halo_info()
            s.set_option("no_overlap", 1)
            s.hxv_device(d_v, d_hv)
            s.sync()
            H.barrier()
            ser = s.time_hxv_passes(d_v, d_hv, max(3, min(steps, 10)))
            s.set_option("no_overlap", 0)
            push = [m for n_, m in ser if n_ == "k_halo_push"]
            res["halo"] = {"bytes_out_per_gpu": H.maxr(float(bo)), "bytes_in_per_gpu": H.maxr(float(bi)), "windows": nwin,
                           "serialised_passes": [(n_, H.maxr(m)) for n_, m in ser],
                           "push_ms": H.maxr(push[0]) if push else None}
            if push and push[0] > 0:
                res["halo"]["nvlink_out_gbs_per_gpu"] = bo / (push[0] * 1e-3) / 1e9
        except Exception as e:                                      # keep the line: the halo block is explanatory
            res["halo"] = {"error": str(e)}
            s.set_option("no_overlap", 0)
        H.barrier()

    # ---- correctness of what was just timed -----------------------------------------------------------
    s.hxv_device(d_v, d_hv)
    s.sync()
    res["parity_check"] = parity_check(H, s, cfg, d_v, d_hv, hermiticity=full)
    H.barrier()

    # ---- Lanczos iterations per second (device-resident recurrence, no host sync inside) ------------
    nl = max(2, min(steps, 50))
    s.time_lanczos_device(d_v, 3)
    H.barrier()
    res["lanczos_ms_per_iter"] = H.maxr(s.time_lanczos_device(d_v, nl)) / nl
    H.barrier()
    if full:
        torch = H.torch
        # ---- end to end through the reference-facing call spHtimesV_p(Nloc, v, Hv): host buffers -------
        hv_h = torch.empty(nloc, dtype=torch.float64).pin_memory()
        v_h = torch.empty(nloc, dtype=torch.float64).pin_memory()
        v_h.numpy()[:] = configs.bench_vector(nloc, s.ishift)
        ke = max(1, min(steps, 8))
        s.spHtimesV_ptr(nloc, v_h.data_ptr(), hv_h.data_ptr())
        H.barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            s.spHtimesV_ptr(nloc, v_h.data_ptr(), hv_h.data_ptr())  # H2D + H*v + D2H, synchronous
        H.barrier()
        res["e2e_s"] = H.maxr((time.perf_counter() - t0) / ke)
        # e2e of a whole chain: one sp_lanc_tridiag call (H2D of the start vector once, coefficients back)
        H.barrier()
        t0 = time.perf_counter()
        s.sp_lanc_tridiag(v_h.numpy(), nl)
        H.barrier()
        res["chain_s"] = H.maxr(time.perf_counter() - t0)
        res["chain_steps"] = nl
        # e2e of the ground-state solve ED_DIAG.f90:174-186 makes: one sp_lanc_eigh call on host arrays (start vector
        # in, eigenvector out, both sweeps of the recurrence on the device), the reference's defaults except that the
        # step count is capped so that the line stays short
        H.barrier()
        nit = 120
        v_h.numpy()[:] = 1.0 / np.sqrt(float(res["dim"]))
        H.barrier()
        t0 = time.perf_counter()
        e0, nle = s.sp_lanc_eigh_ptr(v_h.data_ptr(), nitermax=nit, threshold=1e-12)
        H.barrier()
        res["eigh"] = {"seconds": H.maxr(time.perf_counter() - t0), "nlanc": int(nle), "hxv": 2 * int(nle), "e0": float(e0),
                       "call": "edgpu_sp_lanc_eigh(Nitermax=%d, threshold=1e-12), pinned host vector in/out" % nit}
        del hv_h, v_h
    s.dev_free(d_v)
    s.dev_free(d_hv)
    s.delete_Hv_sector()
    s.close()
    return res


def kernel_table(res, stored, peak):
    bytes_per_el = 24 if stored else 16
    alg = {"k_srow": 16 + (8 if stored else 0), "k_fcol": 24, "k_tile_col": 16 + (8 if stored else 0),
           "k_tile_row": 24, "k_hxv_gather": bytes_per_el}
    kernels = []
    passes = res.get("passes") or []
    tot = sum(m for _, m in passes)
    for name, ms_k in passes:
        b = alg.get(name, bytes_per_el)
        kernels.append({"name": name, "ms": ms_k, "share": ms_k / tot if tot else None, "algorithmic_bytes_per_element": b,
                        "achieved_gbs": b * res["nloc"] / (ms_k * 1e-3) / 1e9, "frac": b * res["nloc"] / (ms_k * 1e-3) / 1e9 / peak})
    return kernels


def run_b200(args):
    H = Harness(args)
    H.init()
    world, rank = H.world, H.rank
    clocks = Clocks(H.local) if rank == 0 else None
    res = measure(H, args.workload, args.steps, full=not args.hxv_only)
    clk = clocks.stop() if clocks else None
    if args.hxv_only:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": res["value"], "unit": UNIT, "ms_per_step": res["ms_per_step"], "hxv_only": True,
                              "gpu_launches": res["launches"], "kernels": res["passes"]}))
        return

    bytes_per_el = 24 if args.stored else 16
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    nloc = res["nloc"]
    ms_step = res["ms_per_step"]
    step_gbs = bytes_per_el * nloc / (ms_step * 1e-3) / 1e9      # per GPU: algorithmic bytes of the local shard
    kernels = kernel_table(res, args.stored, peak)
    dom = max([k for k in kernels if not k["name"].startswith("k_halo")], key=lambda k: k["ms"]) if kernels else None
    if dom:
        traffic, tsrc = dram_traffic(args.workload, args.stored, dom["name"]) if world == 1 else (None, None)
        roof = {"bound": "hbm", "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": dom["frac"], "traffic": traffic,
                "traffic_source": tsrc, "kernel": dom["name"], "ms_per_launch": dom["ms"],
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
        "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload], "operator": OPERATOR,
                   "algo": args.algo, "l2": "input vector %.0f MB per GPU, larger than the 126 MB L2; no flush" % (8 * nloc / 1e6),
                   "sharding": "i_dw columns, %d rank(s)" % world, "vector": "v_i = sin(0.37 i) + 0.1"},
        # the same operator in both arms (stored == direct to 1e-12, tests/test_gpu_parity.py); how each arm evaluates it:
        "operator_form": "stored (ED_SPARSE_H=T)" if args.stored else "direct (ED_SPARSE_H=F, diagonal recomputed)",
        "lanczos_iter_per_s": 1000.0 / res["lanczos_ms_per_iter"],
        "roofline": roof,
        "e2e": {"value": 1.0 / res["e2e_s"], "unit": UNIT, "h2d_bytes_per_step": 8 * res["dim"], "d2h_bytes_per_step": 8 * res["dim"],
                "call": "edgpu_hxv == spHtimesV_p(Nloc,v,Hv) with pinned host arrays",
                "host_gbs_per_gpu_each_way": 8 * nloc / max((res["e2e_s"] - ms_step * 1e-3) / 2, 1e-9) / 1e9,
                "numa_node": H.numa,
                "chain_hxv_per_s": res["chain_steps"] / res["chain_s"],
                "chain_call": "edgpu_sp_lanc_tridiag, %d steps, start vector from host" % res["chain_steps"],
                "eigh": res.get("eigh")},
        "parity_check": res["parity_check"],
        "gpu_launches": res["launches"],
        "clocks": clk,
    }
    if res.get("halo"):
        line["halo"] = res["halo"]
    ok = res["parity_check"]["ok"]
    # ---- the fitted-bath variant (level-dependent V_k: the value-table column kernel instead of the uniform one) ----
    if args.workload == "C3" and not args.no_fitted and not args.stored:
        try:
            r2 = measure(H, "C3V", max(3, min(args.steps, 10)), full=False)
            line["fitted_bath"] = {"workload": WORKLOADS["C3V"], "ms_per_step": r2["ms_per_step"], "hxv_per_s": r2["value"],
                                   "lanczos_iter_per_s": 1000.0 / r2["lanczos_ms_per_iter"],
                                   "kernels": [[n, m] for n, m in (r2.get("passes") or [])], "parity_check": r2["parity_check"]}
            ok = ok and r2["parity_check"]["ok"]
        except Exception as e:
            line["fitted_bath"] = {"error": repr(e)}
    # ---- C5: the Ns=18 sector on the same GPUs (north_star's target configuration) ----------------------
    if args.workload == "C3" and not args.no_c5 and not args.stored:
        try:
            r5 = measure(H, "C5", max(2, min(args.steps, 5)), full=False)
            gbs5 = 16.0 * r5["nloc"] / (r5["ms_per_step"] * 1e-3) / 1e9
            line["c5"] = {"workload": WORKLOADS["C5"], "ms_per_step": r5["ms_per_step"], "hxv_per_s": r5["value"],
                          "lanczos_iter_per_s": 1000.0 / r5["lanczos_ms_per_iter"], "elements_per_gpu": r5["nloc"],
                          "roofline_frac_16B": gbs5 / peak, "kernels": [[n, m] for n, m in (r5.get("passes") or [])],
                          "build_s": r5["build_s"], "parity_check": r5["parity_check"], "halo": r5.get("halo")}
            ok = ok and r5["parity_check"]["ok"]
        except Exception as e:
            line["c5"] = {"error": repr(e)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline(args.workload, args.cpu_seconds)
        except Exception as e:                                   # the baseline never blocks the GPU line
            line["cpu_baseline"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line))
    if H.dist is not None:
        H.dist.destroy_process_group()
    if not ok:
        sys.stderr.write("bench.py: PARITY CHECK FAILED\n")
        sys.exit(3)


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
