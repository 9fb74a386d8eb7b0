"""N>1 host logic on CPU: world_size 2 and 3 over gloo.  Each rank drives the exact send/receive
plan the CUDA path uses (edgpu_transpose_plan / edgpu_split), performs the exchange with
torch.distributed, and the sharded H*v assembled this way must equal the oracle's
spMatVec_MPI_main emulation and the serial operator."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _exchange(send, plan):
    soff, scnt, roff, rcnt = plan
    world, rank = dist.get_world_size(), dist.get_rank()
    recv = torch.zeros(int(roff[-1] + rcnt[-1]), dtype=torch.float64)
    reqs = []
    for p in range(world):
        if p == rank:
            recv[roff[p]:roff[p] + rcnt[p]] = send[soff[p]:soff[p] + scnt[p]]
            continue
        reqs.append(dist.isend(send[soff[p]:soff[p] + scnt[p]].contiguous(), p))
    for p in range(world):
        if p != rank:
            buf = torch.zeros(int(rcnt[p]), dtype=torch.float64)
            dist.recv(buf, p)
            recv[roff[p]:roff[p] + rcnt[p]] = buf
    for r in reqs:
        r.wait()
    return recv


def _worker(rank, world, port, name, sec, q):
    for p in (os.path.join(ROOT, "dmft-lanc-ed_b200"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import edgpu
    import oracle as O
    from edgpu import configs
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = configs.config(name)
        o = O.Oracle(**configs.solver_kwargs(cfg))
        nup, ndw = sec
        with o.sector(nup, ndw) as full, o.sector(nup, ndw, rank, world) as mine:
            du, dd = full.dimup, full.dimdw
            qdw, coloff = edgpu.split(dd, world, rank)
            qup, rowoff = edgpu.split(du, world, rank)
            assert (qdw, coloff * du) == (mine.qdw, mine.ishift)
            v = configs.bench_vector(full.dim)
            ref = full.spmatvec(v)
            V = v[mine.ishift:mine.ishift + mine.nloc].reshape((du, qdw), order="F")
            # forward transpose with the CUDA path's plan and block layouts
            plan = edgpu.transpose_plan(du, dd, world, rank, 0)
            send = torch.zeros(du * qdw, dtype=torch.float64)
            for p in range(world):
                qr, ro = edgpu.split(du, world, p)
                blk = V[ro:ro + qr, :]                               # rows(p) x my columns
                send[plan[0][p]:plan[0][p] + plan[1][p]] = torch.from_numpy(np.ascontiguousarray(blk).ravel())  # [c + qdw*r]
            recv = _exchange(send, plan).numpy()
            Vt = np.zeros((dd, qup))
            for p in range(world):
                qc, co = edgpu.split(dd, world, p)
                Vt[co:co + qc, :] = recv[plan[2][p]:plan[2][p] + plan[3][p]].reshape((qc, qup), order="F")
            Vfull = v.reshape((du, dd), order="F")
            assert np.array_equal(Vt, Vfull[rowoff:rowoff + qup, :].T)       # Appendix B layout
            # H_dw on the transposed shard, then the way back
            rp, cols, vals = mine.hdw()
            Hvt = np.zeros_like(Vt)
            for i in range(dd):
                for k in range(rp[i], rp[i + 1]):
                    Hvt[i, :] += vals[k] * Vt[cols[k], :]
            plan = edgpu.transpose_plan(du, dd, world, rank, 1)
            send = torch.zeros(dd * qup, dtype=torch.float64)
            for p in range(world):
                qc, co = edgpu.split(dd, world, p)
                blk = Hvt[co:co + qc, :]                             # dw-rows(p) x my up-rows
                send[plan[0][p]:plan[0][p] + plan[1][p]] = torch.from_numpy(np.ascontiguousarray(blk).ravel())  # [r + qup*c]
            recv = _exchange(send, plan).numpy()
            HV = np.zeros((du, qdw))
            for p in range(world):
                qr, ro = edgpu.split(du, world, p)
                HV[ro:ro + qr, :] += recv[plan[2][p]:plan[2][p] + plan[3][p]].reshape((qr, qdw), order="F")
            # local part: diagonal + H_up
            rpu, colsu, valsu = mine.hup()
            HV += mine.h0d().reshape((du, qdw), order="F") * V
            for i in range(du):
                for k in range(rpu[i], rpu[i + 1]):
                    HV[i, :] += valsu[k] * V[colsu[k], :]
            got = HV.ravel(order="F")
            sl = slice(mine.ishift, mine.ishift + mine.nloc)
            err = np.abs(got - ref[sl]).max() / np.abs(ref).max()
            assert err < 1e-13, err
            mpi = o.spmatvec_mpi(nup, ndw, world, v)
            assert np.abs(got - mpi[sl]).max() / np.abs(ref).max() < 1e-13
            # Lanczos scalars: all-reduced partial dots equal the global dot
            t = torch.tensor([float(got @ v[sl])], dtype=torch.float64)
            dist.all_reduce(t)
            assert abs(t.item() - ref @ v) < 1e-10 * abs(ref @ v)
        q.put((rank, "ok"))
    except Exception as e:                                           # pragma: no cover
        import traceback
        q.put((rank, "FAIL " + repr(e) + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,name,sec,port", [(2, "C1", (4, 4), 29611), (3, "C1", (5, 3), 29612), (2, "NS6", (3, 2), 29613)])
def test_sharded_operator_over_gloo(world, name, sec, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, sec, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def _halo_worker(rank, world, port, name, ndw, nwin, q):
    """The pushed halo of the sharded fast path, end to end over a real process group: every rank takes its own tables
    from the engine's host code (the triples it has to store, the slot of each of its remote list entries), SENDS the
    listed columns of its shard to the ranks that need them -- window by window, exactly the traffic k_halo_push
    generates -- and every receiver checks that slot s of its halo buffer holds the column its list entry s names."""
    import ctypes as C
    for p in (os.path.join(ROOT, "dmft-lanc-ed_b200"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import edgpu
    from edgpu import configs
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = configs.config(name)
        sv = edgpu.Solver.__new__(edgpu.Solver)
        sv.norb, sv.nbath, sv.nspin = cfg["norb"], cfg["nbath"], cfg["nspin"]
        keep = sv._pack(cfg["uloc"], cfg["ust"], cfg["jh"], cfg["jx"], cfg["jp"], cfg["xmu"], cfg["hfmode"], cfg.get("imphloc"),
                        cfg["bath_e"], cfg["bath_v"], True)
        L = edgpu.selftest_lib()
        import math
        ns = cfg["nbath"] + 1
        dimdw, du = math.comb(ns, ndw), 7                                   # a few rows per column are enough here
        i32p, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
        cap = dimdw * ns + 8
        info = np.zeros(4, np.int32)
        lcol2 = np.zeros(cap, np.int32); pdst = np.zeros(cap, np.int32); pslot = np.zeros(cap, np.int32); psrc = np.zeros(cap, np.int32)
        pwin = np.zeros(nwin + 1, np.int32)
        rc = L.edgpu_selftest_halo_rank(C.byref(keep[0]), ndw, world, rank, C.c_int64(0), C.c_int64(0), nwin, info.ctypes.data_as(i32p),
                                        lcol2.ctypes.data_as(i32p), cap, pdst.ctypes.data_as(i32p), pslot.ctypes.data_as(i32p),
                                        psrc.ctypes.data_as(i32p), cap, pwin.ctypes.data_as(i32p))
        assert rc == 0, rc
        nslot, maxslot, npush, nlist = (int(v) for v in info)
        # my list entries (owner, column in the owner's shard) from the row-kernel plan export
        pinfo = np.zeros(8, np.int32)
        lptr = np.zeros(dimdw + 9, np.int32); lflag = np.zeros(dimdw + 8, np.int32)
        lown = np.zeros(cap, np.int32); lcol = np.zeros(cap, np.int32); lamp = np.zeros(cap)
        rc = L.edgpu_selftest_srow_plan(C.byref(keep[0]), ndw, world, rank, C.c_int64(0), C.c_int64(0), pinfo.ctypes.data_as(i32p),
                                        None, 0, None, 0, None, 0, lptr.ctypes.data_as(i32p), lflag.ctypes.data_as(i32p), dimdw + 8,
                                        lown.ctypes.data_as(i32p), lcol.ctypes.data_as(i32p), lamp.ctypes.data_as(dp), cap)
        assert rc == 0 and pinfo[0] == 1 and pinfo[6] == nlist
        # every rank's slab has the same layout
        ms = torch.tensor([maxslot], dtype=torch.int64)
        allms = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(allms, ms)
        assert all(int(t.item()) == maxslot for t in allms) and nslot <= maxslot
        q_, off = edgpu.split(dimdw, world, rank)
        offs = [edgpu.split(dimdw, world, r)[1] for r in range(world)]
        col = lambda g: np.sin(0.37 * (np.arange(du) + g * du + 1)) + 0.1      # column g of the synthetic vector
        # the push, window by window: counts first (what each peer will receive from me), then the columns
        halo = np.full((max(maxslot, 1), du), np.nan)
        for w in range(nwin):
            tr = range(pwin[w], pwin[w + 1])
            cnt = torch.zeros(world, dtype=torch.int64)
            for k in tr:
                cnt[pdst[k]] += 1
            rcnt = torch.zeros(world, dtype=torch.int64)
            dist.all_to_all_single(rcnt, cnt)
            reqs = []
            for p in range(world):
                if p == rank or cnt[p] == 0:
                    continue
                ks = [k for k in tr if pdst[k] == p]
                payload = np.concatenate([np.concatenate([[float(pslot[k])], col(off + psrc[k])]) for k in ks])
                reqs.append(dist.isend(torch.from_numpy(payload), p))
            for p in range(world):
                if p == rank or rcnt[p] == 0:
                    continue
                buf = torch.zeros(int(rcnt[p]) * (du + 1), dtype=torch.float64)
                dist.recv(buf, p)
                for row in buf.numpy().reshape(-1, du + 1):
                    s = int(row[0])
                    assert 0 <= s < nslot and np.isnan(halo[s, 0])               # every slot is stored exactly once
                    halo[s] = row[1:]
            for r in reqs:
                r.wait()
        # the receiver's view: remote entry e reads slot lcol2[e] and must find the owner's column there
        nrem = 0
        for e in range(nlist):
            if lown[e] == rank:
                assert lcol2[e] == lcol[e]
                continue
            nrem += 1
            assert np.array_equal(halo[lcol2[e]], col(offs[lown[e]] + lcol[e]))
        assert nrem == nslot and not np.isnan(halo[:nslot]).any()
        q.put((rank, "ok"))
    except Exception as e:                                           # pragma: no cover
        import traceback
        q.put((rank, "FAIL " + repr(e) + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,name,ndw,nwin,port", [(2, "C1", 4, 1, 29621), (3, "NS10", 5, 4, 29622), (2, "NS12", 6, 2, 29623)])
def test_pushed_halo_over_gloo(world, name, ndw, nwin, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_halo_worker, args=(r, world, port, name, ndw, nwin, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
