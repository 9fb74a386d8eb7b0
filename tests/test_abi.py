"""The C-ABI library loads, exports every symbol include/*.h declares, fails loudly without a GPU,
and its host-evaluable bit logic (hd_funcs.h, shared with the kernels) matches the oracle."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

import edgpu
from edgpu import configs
from conftest import make_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(edgpu_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = edgpu.lib()
    names = _declared("edgpu.h")
    assert len(names) >= 30
    assert sorted(edgpu.ABI_SYMBOLS) == names
    for n in names:
        assert hasattr(L, n), n
    # the test instrumentation lives in its own library; the product library exports none of it
    T = edgpu.selftest_lib()
    for n in _declared("edgpu_selftest.h"):
        assert hasattr(T, n), n
        assert not hasattr(L, n), n


def test_no_torch_or_cxx_types_in_header():
    txt = open(os.path.join(ROOT, "include", "edgpu.h")).read()
    assert 'extern "C"' in txt
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)          # declarations only
    assert "torch" not in txt and "std::" not in txt and "at::" not in txt


def test_split_is_the_reference_rule():
    # ED_HAMILTONIAN.f90:96-110: first mod(n,P) ranks get one extra, contiguous
    for n, P in [(70, 8), (12870, 8), (48620, 8), (3432, 3), (5, 5), (7, 2)]:
        off = 0
        for r in range(P):
            q, o = edgpu.split(n, P, r)
            assert q == n // P + (1 if r < n % P else 0) and o == off
            off += q
        assert off == n


@pytest.mark.skipif(edgpu.device_count() > 0, reason="checks the no-GPU failure path")
def test_fails_loudly_without_gpu():
    cfg = configs.config("C1")
    with pytest.raises(edgpu.EdgpuError) as e:
        edgpu.Solver(**configs.solver_kwargs(cfg))
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)


def _params(cfg, sparse=True):
    s = edgpu.Solver.__new__(edgpu.Solver)
    s.norb, s.nbath, s.nspin = cfg["norb"], cfg["nbath"], cfg["nspin"]
    return s._pack(cfg["uloc"], cfg["ust"], cfg["jh"], cfg["jx"], cfg["jp"], cfg["xmu"], cfg["hfmode"],
                   cfg.get("imphloc"), cfg["bath_e"], cfg["bath_v"], sparse)


@pytest.mark.parametrize("name", ["C1", "C4", "NS10", "NS12"])
def test_kernel_bit_logic_on_host_matches_oracle(name):
    """hd_funcs.h evaluated on the host through the self-test hooks: sector maps, factor CSR and the
    stored diagonal are BIT-EXACT against the oracle; the factorised diagonal agrees to 1e-14."""
    L = edgpu.selftest_lib()
    L.edgpu_selftest_map.restype = C.c_int64
    L.edgpu_selftest_factor.restype = C.c_int64
    L.edgpu_selftest_diag.restype = C.c_double
    over = {}
    if name == "C4":
        h = np.zeros((1, 1, 2, 2))
        h[0, 0, 0, 1] = h[0, 0, 1, 0] = 0.3
        h[0, 0, 0, 0], h[0, 0, 1, 1] = 0.1, -0.2
        over["imphloc"] = h
    cfg, o = make_oracle(name, **over)
    keep = _params(cfg)
    p = keep[0]
    i32p, i64p, dp = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double)
    for nup, ndw in [(cfg["nup"], cfg["ndw"]), (cfg["nup"] + 1, cfg["ndw"] - 2)]:
        with o.sector(nup, ndw) as s:
            mu, md = s.map_up(), s.map_dw()
            for spin, npart, ref, refmap in [(0, nup, s.hup(), mu), (1, ndw, s.hdw(), md)]:
                n = L.edgpu_selftest_map(o.ns, npart, None)
                m = np.zeros(n, np.int32)
                L.edgpu_selftest_map(o.ns, npart, m.ctypes.data_as(i32p))
                assert np.array_equal(m, refmap)
                nnz = L.edgpu_selftest_factor(C.byref(p), spin, npart, None, None, None)
                rp, cc, vv = np.zeros(n + 1, np.int64), np.zeros(nnz, np.int64), np.zeros(nnz)
                L.edgpu_selftest_factor(C.byref(p), spin, npart, rp.ctypes.data_as(i64p), cc.ctypes.data_as(i64p),
                                        vv.ctypes.data_as(dp))
                assert np.array_equal(rp, ref[0]) and np.array_equal(cc, ref[1]) and np.array_equal(vv, ref[2])
            d = s.h0d()
            for i in np.random.default_rng(0).integers(0, s.dim, 500):
                a = L.edgpu_selftest_diag(C.byref(p), C.c_uint32(int(mu[i % s.dimup])), C.c_uint32(int(md[i // s.dimup])), 0)
                b = L.edgpu_selftest_diag(C.byref(p), C.c_uint32(int(mu[i % s.dimup])), C.c_uint32(int(md[i // s.dimup])), 1)
                assert a == d[i]
                assert abs(b - d[i]) < 1e-14 * max(1.0, abs(d[i]))
            if cfg["norb"] > 1:
                rpn, ccn, vvn = s.hnd()
                cu, cd, va = (C.c_uint32 * 50)(), (C.c_uint32 * 50)(), (C.c_double * 50)()
                for i in np.random.default_rng(1).integers(0, s.dim, 300):
                    k = L.edgpu_selftest_nonlocal_row(C.byref(p), C.c_uint32(int(mu[i % s.dimup])),
                                                      C.c_uint32(int(md[i // s.dimup])), cu, cd, va)
                    cols = [int(np.searchsorted(mu, cu[q])) + int(np.searchsorted(md, cd[q])) * s.dimup for q in range(k)]
                    assert list(ccn[rpn[i]:rpn[i + 1]]) == cols
                    assert list(vvn[rpn[i]:rpn[i + 1]]) == [va[q] for q in range(k)]


def test_host_side_gf_accumulation_matches_oracle():
    """edgpu_add_to_lanczos_gf / sigma_normal are host arithmetic on chain outputs: checked on CPU."""
    import oracle as O
    gold = np.load(os.path.join(ROOT, "tests", "golden", "c1_golden.npz"))
    z = 1j * gold["wm"]
    g = edgpu.add_to_lanczos_gf(float(gold["add_norm2"]), float(gold["e0"]), gold["add_alanc"], gold["add_blanc"], 1, z)
    g += edgpu.add_to_lanczos_gf(float(gold["rem_norm2"]), float(gold["e0"]), gold["rem_alanc"], gold["rem_blanc"], -1, z)
    assert np.abs(g - gold["gmats"]).max() < 1e-12
    cfg = configs.config("C1")
    sig, _ = edgpu.sigma_normal(z, g, cfg["xmu"], 0.0, cfg["bath_e"], cfg["bath_v"])
    assert np.abs(sig - gold["smats"]).max() < 1e-10


ROWPLAN_CASES = [("C1", 4, 1, 0, 0), ("C1", 4, 2, 0, 0), ("C1", 4, 3, 0, 1), ("C1", 4, 4, 0, 2), ("C1", 3, 4, 4, 1),
                 ("NS10", 5, 3, 4, 3), ("NS10", 5, 8, 0, 0), ("NS12", 6, 7, 0, 4), ("NS12", 5, 8, 4, 2),
                 ("NS14", 7, 8, 0, 0), ("NS16", 8, 8, 0, 0), ("NS16", 9, 5, 0, 0), ("NS16", 8, 1, 0, 0), ("NS18", 9, 8, 0, 0)]


def _row_plan(p, ndw, nranks, rank, lr, st, dimdw, nnz):
    """edgpu_selftest_srow_plan -> dict of numpy arrays."""
    L = edgpu.selftest_lib()
    i32p, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    info = np.zeros(8, np.int32)
    cap_j, cap_c, cap_r, cap_cols, cap_e = 1 << 16, 1 << 14, 1 << 15, dimdw + 8, nnz + 8
    jhi = np.zeros(cap_j, np.int32); chunks = np.zeros(4 * cap_c, np.int32); recs = np.zeros(20 * cap_r, np.int32)
    lptr = np.zeros(cap_cols + 1, np.int32); lflag = np.zeros(cap_cols, np.int32)
    lown = np.zeros(cap_e, np.int32); lcol = np.zeros(cap_e, np.int32); lamp = np.zeros(cap_e)
    rc = L.edgpu_selftest_srow_plan(C.byref(p), ndw, nranks, rank, C.c_int64(lr), C.c_int64(st), info.ctypes.data_as(i32p),
                                    jhi.ctypes.data_as(i32p), cap_j, chunks.ctypes.data_as(i32p), cap_c,
                                    recs.ctypes.data_as(i32p), cap_r, lptr.ctypes.data_as(i32p), lflag.ctypes.data_as(i32p), cap_cols,
                                    lown.ctypes.data_as(i32p), lcol.ctypes.data_as(i32p), lamp.ctypes.data_as(dp), cap_e)
    assert rc == 0 and info[0] == 1, (rc, info)
    return dict(LR=int(info[1]), T=int(info[2]), nhigh=int(info[3]), jhi=jhi[:1 << int(info[3])],
                chunks=chunks[:4 * info[4]].reshape(-1, 4), recs=recs[:20 * info[5]].reshape(-1, 20).view(np.uint32),
                lptr=lptr, lflag=lflag, lown=lown[:info[6]], lcol=lcol[:info[6]], lamp=lamp[:info[6]], nzcols=int(info[7]))


def _kernel_hops(plan, ndw, vk):
    """What k_srow does with the plan (hxv_fast.cu: srow_group / srow_hops), restated: {(target, source): value} in
    shard-local column indices, and the set of columns it writes."""
    LR, T, nhigh = plan["LR"], plan["T"], plan["nhigh"]
    popc = lambda x: bin(x).count("1")
    pats = {n: [q for q in range(1 << LR) if popc(q) == n] for n in range(LR + 1)}
    rank = lambda lo: pats[popc(lo)].index(lo)
    hops, written = {}, []
    for rb, ng, cb, ce in plan["chunks"]:
        for rec in plan["recs"][rb:rb + ng]:
            lb, N, hx, par = int(rec[0]), int(rec[1]), int(rec[2]), int(rec[3])
            pc = rec[4:].view(np.int32)
            h, ex = hx & 0xFFFF, hx >> 16
            assert N == ndw - popc(h) and 0 <= lb and cb + lb + len(pats[N]) <= ce
            for i, lo in enumerate(pats[N]):
                t = cb + lb + i
                written.append(t)
                for kb in range(1, LR):                               # low bits: register to register
                    if ((lo >> kb) & 1) != (lo & 1):
                        sgn = -1.0 if popc(lo & ((1 << kb) - 2)) & 1 else 1.0
                        hops[(t, cb + lb + rank(lo ^ (1 | (1 << kb))))] = sgn * vk[kb]
                for kk in range(nhigh):                               # high bits: whole group -> one partner group
                    if not (ex >> kk) & 1 or bool(h & (1 << kk)) == bool(lo & 1):
                        continue
                    base2 = (cb if kk < T else 0) + int(pc[kk])
                    sgn = -1.0 if (popc(lo >> 1) + ((par >> kk) & 1)) & 1 else 1.0
                    hops[(t, base2 + rank(lo ^ 1))] = sgn * vk[LR + kk]
    return hops, written


@pytest.mark.parametrize("name,ndw,nranks,lr,st", ROWPLAN_CASES)
def test_row_kernel_plan_covers_every_hop_exactly_once(name, ndw, nranks, lr, st):
    """Host arithmetic behind the structured row kernel and its sharded form (hxv_fast.cu: srow_plan_host /
    srow_lists_host), checked without a GPU for every rank of a split: the records, evaluated the way the kernel
    evaluates them, plus the source lists of the halo kernel reproduce spH0dws(1) of the oracle (stored/H_dw.f90:8-80)
    restricted to the rank's target columns EXACTLY once per entry, values bit-exact; every local column is written
    by exactly one of the two passes; list entries name valid owner columns."""
    cfg, o = make_oracle(name)
    keep = _params(cfg)
    ns = cfg["nbath"] + 1
    dimdw = math.comb(ns, ndw)
    big = dimdw > 20000                                           # Ns = 18: check against the closed-form structure
    if not big:
        with o.sector(cfg["nup"], ndw) as s:
            rp, cc, vv = s.hdw()
            md = s.map_dw()
        ref_all = {(t, int(cc[e])): float(vv[e]) for t in range(dimdw) for e in range(rp[t], rp[t + 1])}
        nnz = len(cc)
    else:
        nnz = dimdw * ns
    vkall = [0.0] + [float(cfg["bath_v"].reshape(-1)[k]) for k in range(ns - 1)]
    total = 0
    for rank in range(nranks):
        plan = _row_plan(keep[0], ndw, nranks, rank, lr, st, dimdw, nnz)
        q, off = edgpu.split(dimdw, nranks, rank)
        coloffs = [edgpu.split(dimdw, nranks, r)[1] for r in range(nranks)] + [dimdw]
        hops, written = _kernel_hops(plan, ndw, vkall)
        assert len(written) == len(set(written))
        got = {}
        for (t, s_), v in hops.items():
            assert 0 <= t < q and 0 <= s_ < q                      # the row kernel never leaves the shard
            got[(off + t, off + s_)] = v
        cover = np.zeros(q, np.int32)
        cover[written] += 1
        if nranks == 1:
            assert len(plan["lown"]) == 0 and plan["nzcols"] == 0
        else:
            lptr, lflag = plan["lptr"], plan["lflag"]
            cover += lflag[:q] & 1
            nz = 0
            for t in range(q):
                assert bool(lflag[t] & 2) == (lptr[t + 1] > lptr[t])
                nz += int(lptr[t + 1] > lptr[t])
                for e in range(lptr[t], lptr[t + 1]):
                    own = int(plan["lown"][e])
                    src = coloffs[own] + int(plan["lcol"][e])
                    assert 0 <= own < nranks and coloffs[own] <= src < coloffs[own + 1]
                    assert (off + t, src) not in got
                    got[(off + t, src)] = float(plan["lamp"][e])
            assert nz == plan["nzcols"]
        assert (cover == 1).all()                                  # every local column written by exactly one kernel
        got = {k: v for k, v in got.items() if v != 0.0}           # the reference stores no entry for V_k = 0
        if not big:
            ref = {k: v for k, v in ref_all.items() if off <= k[0] < off + q}
            assert got.keys() == ref.keys()
            assert all(got[k] == ref[k] for k in ref)
        else:
            assert all(off <= k[0] < off + q for k in got)
        total += len(got)
    if big:                                                        # star geometry at half filling: exactly Ns/2 hops per column
        assert total == dimdw * ns // 2 or ndw * 2 != ns
    # chunk geometry: tiles fit the box grid and the group count the kernel assumes
    assert plan["T"] <= 5 and all(ng <= 64 for _, ng, _, _ in plan["chunks"])


HALO_CASES = [("C1", 4, 2, 0, 0, 1), ("C1", 4, 3, 0, 1, 2), ("C1", 3, 4, 4, 1, 4), ("NS10", 5, 8, 0, 0, 4), ("NS12", 6, 7, 0, 4, 3),
              ("NS12", 5, 8, 4, 2, 8), ("NS14", 7, 8, 0, 0, 4), ("NS16", 8, 2, 0, 0, 1), ("NS16", 8, 8, 0, 0, 4), ("NS16", 9, 5, 0, 0, 4),
              ("NS18", 9, 8, 0, 0, 4)]


@pytest.mark.parametrize("name,ndw,nranks,lr,st,nwin", HALO_CASES)
def test_halo_push_tables_match_the_source_lists(name, ndw, nranks, lr, st, nwin):
    """Host arithmetic of the sharded fast path's halo (hxv_fast.cu: halo_tables_host), without a GPU and for every rank
    at once: what the owners store into their peers' halo buffers (k_halo_push) is exactly what the column pass of the
    receivers reads -- one (peer, slot, source column) triple per remote list entry, filed under the column window of
    its target, none left over, identical slab layout on every rank.  Replaces the per-column MPI_AllToAllV of
    vector_transpose_MPI (ED_HAMILTONIAN_COMMON.f90:53-118) on this path."""
    cfg, _ = make_oracle(name)
    keep = _params(cfg)
    L = edgpu.selftest_lib()
    info = np.zeros(6, np.int32)
    rc = L.edgpu_selftest_halo_tables(C.byref(keep[0]), ndw, nranks, C.c_int64(lr), C.c_int64(st), nwin,
                                      info.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == 0
    ok, entries, remote, triples, maxslot, k = (int(v) for v in info)
    assert ok == 1, info
    assert remote == triples and entries >= remote and k == nwin
    assert remote > 0 and maxslot > 0                               # every split of these sectors crosses a rank boundary


@pytest.mark.parametrize("name,nups,ndws", [("ORB2", [1, 2], [2, 1]), ("ORB2B", [3, 2], [2, 3]), ("ORB3", [2, 1, 2], [1, 2, 1])])
def test_orbs_factors_on_host_match_oracle(name, nups, ndws):
    """Host arithmetic of the orbital-resolved sectors (csrc/orbs.cu: orbs_factor_host), without a GPU: every word map
    and every spH0ups(iorb) / spH0dws(iorb) -- structure in insertion order AND values -- bit-exact against the oracle's
    ed_buildh_orbs restatement (ED_HAMILTONIAN_SPARSE_HxV.f90:206-370, stored/Orbs/H_up.f90, H_dw.f90)."""
    cfg, o = make_oracle(name)
    keep = _params(cfg)
    L = edgpu.selftest_lib()
    L.edgpu_selftest_orbs_factor.restype = C.c_int64
    i64p = C.POINTER(C.c_int64)
    with o.sector_orbs(nups, ndws) as so:
        for f in range(2 * cfg["norb"]):
            n = (nups + ndws)[f]
            nnz = C.c_int64(0)
            dim = L.edgpu_selftest_orbs_factor(C.byref(keep[0]), f, n, None, C.byref(nnz), None, None, None)
            assert dim == so.dims[f]
            m = np.zeros(dim, np.int32); rp = np.zeros(dim + 1, np.int64)
            cols = np.zeros(max(nnz.value, 1), np.int64); vals = np.zeros(max(nnz.value, 1))
            L.edgpu_selftest_orbs_factor(C.byref(keep[0]), f, n, m.ctypes.data_as(C.POINTER(C.c_int32)), None, rp.ctypes.data_as(i64p),
                                         cols.ctypes.data_as(i64p), vals.ctypes.data_as(C.POINTER(C.c_double)))
            om, orp, ocols, ovals = so.factor(f)
            assert np.array_equal(m, om) and np.array_equal(rp, orp)
            assert np.array_equal(cols[:nnz.value], ocols) and np.array_equal(vals[:nnz.value], ovals)


def test_header_is_plain_c_and_the_c_host_example_links(tmp_path):
    """include/edgpu.h is a C header (the Fortran shim binds it through ISO_C_BINDING): integration/example_host.c, a
    plain-C99 host that drives the reference's call sequence (build_Hv_sector, spHtimesV_p, sp_lanc_eigh, the GF chains,
    the observables), compiles with -Wall -Wextra and links against libedgpu.so.  (Running it needs a B200.)"""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = os.path.join(ROOT, "integration", "example_host.c")
    inc = os.path.join(ROOT, "include")
    libdir = os.path.join(ROOT, "dmft-lanc-ed_b200")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", inc, "-fsyntax-only", src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = str(tmp_path / "example_host")
    r = subprocess.run([gcc, "-std=c99", "-I", inc, src, "-L", libdir, "-ledgpu", "-Wl,--allow-shlib-undefined", "-lm", "-o", out],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_add_to_lanczos_chi_host_formula_on_cpu():
    """edgpu_add_to_lanczos_chi is host arithmetic on a chain's coefficients (add_to_lanczos_spinChi / _densChi,
    ED_GF_CHISPIN.f90:434-488): checked on the CPU against the oracle's restatement on real chain coefficients, and on a
    one-pole chain against the closed form chi(tau) = norm2 exp(-tau dE), chi(i nu) = norm2 (1 - e^{-beta dE}) 2 dE / (nu^2 + dE^2)."""
    import oracle as O
    cfg, o = make_oracle("NS6")
    with o.sector(3, 3) as s:
        e0, gs, _, _ = s.lanc_eigh(v0=np.ones(s.dim) / np.sqrt(s.dim))
    beta = 40.0
    vm = np.pi / beta * 2 * np.arange(7)
    tau = np.linspace(0.0, beta, 9)
    vr = np.linspace(-2.5, 2.5, 11)
    for kind in (0, 1):
        ch = o.chi_chain(3, 3, gs, kind, 1, 1, ngfiter=40)
        got = edgpu.add_to_lanczos_chi(ch["norm2"], e0, beta, ch["alanc"], ch["blanc"], vm, tau, vr, 0.02)
        want = O.add_to_lanczos_chi(ch["norm2"], e0, beta, ch["alanc"], ch["blanc"], vm, tau, vr, 0.02)
        for g_, w_ in zip(got, want):
            assert np.abs(g_ - w_).max() < 1e-11
    de, n2 = 0.7, 0.3
    civ, ctau, cw = edgpu.add_to_lanczos_chi(n2, -1.0, beta, np.array([-1.0 + de]), np.array([0.0]), vm, tau, vr, 0.02)
    bose = 1.0 - np.exp(-beta * de)
    assert np.abs(ctau - n2 * np.exp(-tau * de)).max() < 1e-14
    assert abs(civ[0] - n2 * 2 * bose / de) < 1e-14
    assert np.abs(civ[1:] - n2 * bose * 2 * de / (vm[1:] ** 2 + de ** 2)).max() < 1e-14
    w = vr + 0.02j
    assert np.abs(cw + n2 * bose * (1.0 / (w - de) - 1.0 / (w + de))).max() < 1e-13
