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
