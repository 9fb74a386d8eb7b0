"""Launched under torchrun (one process per GPU) by tests/test_gpu_multi.py or by hand:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_worker.py
Checks the sharded operator, Lanczos and GF chains against the oracle (rank 0 holds the reference)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "dmft-lanc-ed_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import edgpu  # noqa: E402
import oracle as O  # noqa: E402
from edgpu import configs  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    def fresh_uid():
        """A NCCL unique id serves ONE communicator: every Solver gets its own (rank 0 draws, broadcast)."""
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(edgpu.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        return bytes(idt.cpu().numpy().tobytes())

    fails = []

    def check(name, ok, detail=""):
        if not ok:
            fails.append("%s rank %d %s" % (name, rank, detail))

    # (config, sector, stored?, engine options).  Single-band cases run the sharded fast path (halo columns pushed by
    # their owners over NVLink next to k_srow, remote / boundary hops applied by the column pass); "no_peer" forces the
    # all-to-all transposes; C4 (spin-exchange / pair-hopping) uses the all-gather path.
    cases = [("C1", (4, 4), True, {}), ("C1", (4, 4), False, {}), ("C1", (5, 3), True, {"srow_t": 1}),
             ("NS10", (5, 5), False, {"srow_lr": 4, "srow_t": 3}), ("NS10", (6, 4), True, {}),
             ("NS12", (6, 6), False, {}), ("NS12", (7, 5), True, {"srow_t": 4}), ("NS10V", (5, 5), False, {}),
             ("NS10", (5, 5), True, {"no_peer": 1}), ("NS12", (6, 6), False, {"no_peer": 1, "hxv_algo": 1}),
             ("C4", (5, 5), True, {}), ("C4", (5, 5), False, {}), ("C4", (6, 5), True, {})]
    if os.environ.get("EDGPU_WORKER_QUICK"):                # a short subset (8-rank runs: every rank repeats the CPU oracle)
        cases = [cases[k] for k in (0, 3, 6, 7, 8, 10)]
    for name, sec, sparse, opts in cases:
        cfg = configs.config(name)
        kw = configs.solver_kwargs(cfg)
        o = O.Oracle(**kw)
        s = edgpu.Solver(ed_sparse_h=sparse, device=local, **kw)
        for k_, v_ in opts.items():
            s.set_option(k_, v_)
        s.set_comm(rank, world, fresh_uid())
        with o.sector(*sec) as full, o.sector(sec[0], sec[1], rank, world) as mine:
            v = configs.bench_vector(full.dim)
            v /= np.linalg.norm(v)
            ref = full.spmatvec(v)
            isec = s.get_sector(*sec)
            check("vecdim", s.vecDim_Hv_sector(isec) == mine.nloc)
            s.build_Hv_sector(isec)
            check("ishift", s.ishift == mine.ishift and s.nloc == mine.nloc)
            sl = slice(mine.ishift, mine.ishift + mine.nloc)
            if sparse:
                check("diag", np.array_equal(s.diag(), mine.h0d()))
                rp, cols, vals = s.csr(2)
                orp, ocols, ovals = mine.hnd()
                if cfg["norb"] > 1:
                    check("hnd", np.array_equal(rp, orp) and np.array_equal(cols, ocols) and np.array_equal(vals, ovals))
            hv = s.spHtimesV(v[sl])
            err = np.abs(hv - ref[sl]).max() / np.abs(ref).max()
            check("hxv %s %s sparse=%s" % (name, sec, sparse), err < 1e-13, "err %.3e" % err)
            # Lanczos ground state on shards
            v0 = np.ones(full.dim) / np.sqrt(full.dim)
            e_ref, vec_ref, a_ref, b_ref = full.lanc_eigh(v0=v0)
            e0, vec, a, b = s.sp_lanc_eigh(v0[sl])
            check("E0", abs(e0 - e_ref) < 1e-12 * abs(e_ref), "%.15g vs %.15g" % (e0, e_ref))
            n = min(20, len(a), len(a_ref))          # well inside the numerically stable prefix (tests/test_gpu_parity.py)
            check("alanc", np.abs(a[:n] - a_ref[:n]).max() < 1e-8 and np.abs(b[:n] - b_ref[:n]).max() < 1e-8)
            # same eigenvector up to a global sign: overlap of the sharded vector with the oracle's (second order in
            # the vector error, like the single-GPU test)
            ov = torch.tensor([float(vec @ vec_ref[sl]), float(vec @ vec)], dtype=torch.float64, device="cuda")
            dist.all_reduce(ov)
            check("vec", abs(abs(ov[0].item()) - 1) < 1e-10 and abs(ov[1].item() - 1) < 1e-12, "overlap %.3e" % (abs(ov[0].item()) - 1))
            s.delete_Hv_sector()
            # GF chains from the sharded ground state: spin-up channels act inside the local columns, spin-down
            # channels are a signed column permutation (one grouped exchange between the ranks)
            s.gf_set_state(isec, vec_ref[sl], e_ref)
            chans = [(io, sp, sg) for io in range(1, cfg["norb"] + 1) for sp in (1, 2) for sg in (1, -1)]
            res = s.gf_chains(chans, nlanc_max=60)
            k = 0
            for io in range(1, cfg["norb"] + 1):
                for sp in (1, 2):
                    rr = o.build_gf_normal(sec[0], sec[1], vec_ref, e_ref, io, ispin=sp, ngfiter=60, lmats=8, lreal=8)
                    for p in range(2):
                        r, rc = res[k], rr["chains"][p]
                        k += 1
                        check("gf nlanc", r["nlanc"] == rc["nlanc"], "%d vs %d" % (r["nlanc"], rc["nlanc"]))
                        if rc["nlanc"]:
                            m = min(20, rc["nlanc"])
                            check("gf norm2 spin %d" % sp, abs(r["norm2"] - rc["norm2"]) < 1e-12)
                            check("gf a spin %d" % sp, np.abs(r["alanc"][:m] - rc["alanc"][:m]).max() < 1e-8)
                            check("gf b spin %d" % sp, np.abs(r["blanc"][:m] - rc["blanc"][:m]).max() < 1e-8)
            # susceptibility chains on shards: a diagonal operator on every rank's own columns (ED_GF_CHISPIN.f90:114-207)
            if cfg["norb"] == 1:
                rs = s.chi_chains(0, [(1, 1), (0, 0)], nlanc_max=40)
                rc_ = o.chi_chain(sec[0], sec[1], vec_ref, 0, 1, 1, ngfiter=40)
                for r in rs:
                    m = min(15, rc_["nlanc"])
                    check("chi norm2", abs(r["norm2"] - rc_["norm2"]) < 1e-12)
                    check("chi a", r["nlanc"] == rc_["nlanc"] and np.abs(r["alanc"][:m] - rc_["alanc"][:m]).max() < 1e-8)
        s.close()
    # one phonon mode on shards: every rank holds DimPh slabs of its own i_dw columns (v(i + (iph-1)*DimUp*mpiQdw))
    for name, sec, nph in [("NS10", (5, 5), 2)]:
        cfg = configs.config(name)
        kw = configs.solver_kwargs(cfg)
        o = O.Oracle(**kw)
        s = edgpu.Solver(nph=nph, g_ph=(0.4,), w0_ph=0.6, device=local, **kw)
        s.set_comm(rank, world, fresh_uid())
        with o.sector(*sec) as full, o.sector(sec[0], sec[1], rank, world) as mine:
            n = full.dim * (nph + 1)
            v = configs.bench_vector(n)
            v /= np.linalg.norm(v)
            ref = full.spmatvec_ph(v, nph, (0.4,), 0.6)
            s.build_Hv_sector(s.get_sector(*sec))
            check("ph nloc", s.nloc == mine.nloc * (nph + 1))
            loc = np.concatenate([v[k * full.dim + mine.ishift:k * full.dim + mine.ishift + mine.nloc] for k in range(nph + 1)])
            rloc = np.concatenate([ref[k * full.dim + mine.ishift:k * full.dim + mine.ishift + mine.nloc] for k in range(nph + 1)])
            hv = s.spHtimesV(loc)
            err = np.abs(hv - rloc).max() / np.abs(ref).max()
            check("ph hxv", err < 1e-13, "err %.3e" % err)
            v0 = np.ones(n) / np.sqrt(n)
            e_ref = full.lanc_eigh_ph(nph, (0.4,), 0.6, v0)[0]
            l0 = np.concatenate([v0[k * full.dim + mine.ishift:k * full.dim + mine.ishift + mine.nloc] for k in range(nph + 1)])
            e0 = s.sp_lanc_eigh(l0)[0]
            check("ph E0", abs(e0 - e_ref) < 1e-12 * abs(e_ref), "%.15g vs %.15g" % (e0, e_ref))
            s.delete_Hv_sector()
        s.close()
    t = torch.tensor([len(fails)], device="cuda")
    dist.all_reduce(t)
    for f in fails:
        print("FAIL:", f, flush=True)
    if rank == 0:
        print("MULTIGPU %s: %d ranks, %d failures" % ("OK" if t.item() == 0 else "FAILED", world, int(t.item())), flush=True)
    dist.destroy_process_group()
    sys.exit(1 if t.item() else 0)


if __name__ == "__main__":
    main()
