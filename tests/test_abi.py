"""The C-ABI library loads, exports every symbol include/*.h declares, fails loudly without a GPU,
and its host-evaluable bit logic (hd_funcs.h, shared with the kernels) matches the oracle."""
import ctypes as C
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
