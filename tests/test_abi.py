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
    for n in names + _declared("edgpu_selftest.h"):
        assert hasattr(L, n), n


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
    L = edgpu.lib()
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


ROWPLAN_CASES = [("C1", 4, 1, 0, 0), ("C1", 4, 2, 0, 0), ("C1", 4, 3, 0, 0), ("C1", 4, 4, 0, 12), ("C1", 3, 4, 4, 12),
                 ("NS10", 5, 3, 4, 24), ("NS10", 5, 8, 0, 0), ("NS12", 6, 7, 0, 64), ("NS12", 5, 8, 4, 48),
                 ("NS14", 7, 8, 0, 0), ("NS16", 8, 8, 0, 0), ("NS16", 9, 5, 0, 0)]


@pytest.mark.parametrize("name,ndw,nranks,lr,cmax", ROWPLAN_CASES)
def test_row_kernel_plan_covers_every_hop_exactly_once(name, ndw, nranks, lr, cmax):
    """Host arithmetic behind the sharded structured row kernel (hxv_fast.cu: srow_plan_host / srow_fix_host),
    checked without a GPU for every rank of a split: the chunks tile the whole low groups of the rank, the
    fix-up list holds exactly the hops whose target or source group is cut by a rank boundary (with the oracle's
    matrix elements), every other hop is left to the kernel, and the Lin table is the closed-form rank."""
    L = edgpu.lib()
    cfg, o = make_oracle(name)
    keep = _params(cfg)
    p = keep[0]
    i32p, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    with o.sector(cfg["nup"], ndw) as s:
        rp, cc, vv = s.hdw()
        md = s.map_dw()
        dimdw = len(md)
    handled = np.zeros(len(cc), dtype=np.int32)          # how many times each CSR entry (target <- source) is covered
    colcov = np.zeros(dimdw, dtype=np.int32)
    for rank in range(nranks):
        info = np.zeros(8, np.int32)
        cap_j, cap_c, cap_t, cap_e = 1 << 15, 4096, dimdw + 8, len(cc) + 8
        jhi = np.zeros(cap_j, np.int32); chunks = np.zeros(4 * cap_c, np.int32)
        tcol = np.zeros(cap_t, np.int32); tinit = np.zeros(cap_t, np.int32); tptr = np.zeros(cap_t + 1, np.int32)
        eown = np.zeros(cap_e, np.int32); esrc = np.zeros(cap_e, np.int32); ev = np.zeros(cap_e)
        rc = L.edgpu_selftest_srow_plan(C.byref(p), ndw, nranks, rank, C.c_int64(lr), C.c_int64(cmax),
                                        info.ctypes.data_as(i32p), jhi.ctypes.data_as(i32p), cap_j,
                                        chunks.ctypes.data_as(i32p), cap_c, tcol.ctypes.data_as(i32p),
                                        tinit.ctypes.data_as(i32p), tptr.ctypes.data_as(i32p), cap_t,
                                        eown.ctypes.data_as(i32p), esrc.ctypes.data_as(i32p), ev.ctypes.data_as(dp), cap_e)
        assert rc == 0 and info[0] == 1, (rc, info)
        LR, nhigh, nchunks, nfix = int(info[1]), int(info[2]), int(info[4]), int(info[6])
        q, off = edgpu.split(dimdw, nranks, rank)
        coloffs = [edgpu.split(dimdw, nranks, r)[1] for r in range(nranks)] + [dimdw]
        # Lin table: first column of every group = rank of its smallest word; owner / cut flags from the split
        lowmask = (1 << LR) - 1
        gcut = {}
        for h in range(1 << nhigh):
            e = int(jhi[h])
            nlow = ndw - bin(h).count("1")
            if nlow < 0 or nlow > LR:
                assert e == -1
                continue
            words = md[(md >> LR) == h]
            base = int(np.searchsorted(md, words[0]))
            assert (e & 0xFFFFF) == base and len(words) == math.comb(LR, nlow)
            own = max(r for r in range(nranks) if coloffs[r] <= base)
            cut = base + len(words) > coloffs[own + 1]
            assert ((e >> 20) & 63) == own and bool(e & 0x40000000) == cut
            gcut[h] = cut
        colcut = np.array([gcut[int(w) >> LR] for w in md])
        # chunks: whole, uncut local groups, contiguous, within the size limit
        ch = chunks[:4 * nchunks].reshape(-1, 4)
        for k in range(nchunks):
            assert off <= ch[k, 2] < ch[k, 3] <= off + q and ch[k, 3] - ch[k, 2] <= info[5]
            assert not colcut[ch[k, 2]:ch[k, 3]].any()
            colcov[ch[k, 2]:ch[k, 3]] += 1
            if k:
                assert ch[k, 2] == ch[k - 1, 3] and ch[k, 0] == ch[k - 1, 1]
        inchunk = np.zeros(dimdw, bool)
        for k in range(nchunks):
            inchunk[ch[k, 2]:ch[k, 3]] = True
        # the kernel's rule: a hop is applied by k_srow iff its target is in a chunk and its source group is not cut
        for t in range(off, off + q):
            for e in range(rp[t], rp[t + 1]):
                if inchunk[t] and not colcut[cc[e]]:
                    handled[e] += 1
        # fix-up list: targets in cut groups are initialised there, every listed edge is a real matrix element
        for k in range(nfix):
            t = off + int(tcol[k])
            assert bool(tinit[k]) == bool(colcut[t])
            if tinit[k]:
                colcov[t] += 1
            for e in range(tptr[k], tptr[k + 1]):
                src = coloffs[int(eown[e])] + int(esrc[e])
                hit = [x for x in range(rp[t], rp[t + 1]) if cc[x] == src]
                assert len(hit) == 1 and vv[hit[0]] == ev[e]
                handled[hit[0]] += 1
    assert (colcov == 1).all()                                # every column written by exactly one kernel
    assert (handled == 1).all()                               # every hop applied exactly once


@pytest.mark.parametrize("name,ndw,lr", [("C1", 4, 5), ("C1", 5, 4), ("NS10", 5, 5), ("NS10V", 4, 4), ("NS12V", 6, 5), ("NS14", 7, 5)])
def test_structured_hop_enumeration_reproduces_spH0dws(name, ndw, lr):
    """The algebra k_srow relies on (hxv_fast.cu: srow_group / srow_prepare), restated in Python on top of the
    engine's own Lin table: low group = (high word h, LR low bits), hop = impurity bit 0 <-> bath bit k, partner
    column = jhi[h'] + rank of the partner's low pattern in its class, sign = parity of the occupied bits strictly
    between the two.  Enumerated that way, the (target, source, value) triples must be EXACTLY spH0dws(1) of the
    oracle (stored/H_dw.f90:8-80) -- structure bit-exact, values bit-exact."""
    L = edgpu.lib()
    cfg, o = make_oracle(name)
    keep = _params(cfg)
    i32p = C.POINTER(C.c_int32)
    with o.sector(cfg["nup"], ndw) as s:
        rp, cc, vv = s.hdw()
    info = np.zeros(8, np.int32)
    jhi = np.zeros(1 << 15, np.int32)
    rc = L.edgpu_selftest_srow_plan(C.byref(keep[0]), ndw, 1, 0, C.c_int64(lr), C.c_int64(0), info.ctypes.data_as(i32p),
                                    jhi.ctypes.data_as(i32p), len(jhi), None, 0, None, None, None, 0, None, None, None, 0)
    assert rc == 0 and info[0] == 1
    LR, nhigh = int(info[1]), int(info[2])
    ns = LR + nhigh
    vk = [0.0] + [float(cfg["bath_v"].reshape(-1)[k]) for k in range(ns - 1)]      # V_k of the dw spin (Nspin = 1)
    popc = lambda x: bin(x).count("1")
    pats = {n: [q for q in range(1 << LR) if popc(q) == n] for n in range(LR + 1)}
    rank = lambda lo: pats[popc(lo)].index(lo)
    trip = {}
    for h in range(1 << nhigh):
        n = ndw - popc(h)
        if n < 0 or n > LR:
            continue
        base = int(jhi[h]) & 0xFFFFF
        for i, lo in enumerate(pats[n]):
            t = base + i
            for kb in range(1, LR):                                   # hops among the low bits (register to register)
                if ((lo >> kb) & 1) != (lo & 1):
                    lo2 = lo ^ (1 | (1 << kb))
                    sgn = -1.0 if popc(lo & ((1 << kb) - 2)) & 1 else 1.0
                    trip[(t, base + rank(lo2))] = sgn * vk[kb]
            for kk in range(nhigh):                                   # hops on the high bits (whole group -> one group)
                bit = 1 << kk
                if bool(h & bit) == bool(lo & 1):
                    continue                                          # bath bit and impurity bit must differ
                lo2 = lo ^ 1
                base2 = int(jhi[h ^ bit]) & 0xFFFFF
                sgn = -1.0 if (popc(lo >> 1) + popc(h & (bit - 1))) & 1 else 1.0
                trip[(t, base2 + rank(lo2))] = sgn * vk[LR + kk]
    ref = {(t, int(cc[e])): float(vv[e]) for t in range(len(rp) - 1) for e in range(rp[t], rp[t + 1])}
    trip = {k: v for k, v in trip.items() if v != 0.0}                # the reference stores no entry for V_k = 0
    assert trip.keys() == ref.keys()
    assert all(trip[k] == ref[k] for k in ref)
