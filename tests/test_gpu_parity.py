"""Parity of the CUDA path (through the C-ABI, ctypes) against the CPU oracle.  Run with -m gpu.

Tolerances (BASELINE.json north_star): basis maps and CSR structure bit-exact; H*v per element
1e-13 relative to |Hv|_max (summation order differs between the reference's own variants, SURVEY
7.3-8); E0 1e-12 relative; Lanczos a_n/b_n 1e-8 for the first 50 steps; G(iw), Sigma(iw) 1e-8 absolute.
"""
import os

import numpy as np
import pytest

import edgpu
import oracle as O
from edgpu import configs
from conftest import make_oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stable_steps(a1, b1, a2, b2, nmax=50, tol=1e-10):
    """Number of leading Lanczos steps for which the coefficient comparison is well posed.

    The three-term recurrence amplifies rounding differences exponentially once extremal Ritz values
    have converged, so two CORRECT evaluations that differ only in summation order (the reference's
    own stored and direct variants) drift apart after some step.  a1/b1 and a2/b2 are
    the oracle's coefficients from spMatVec_main and directMatVec_main; beyond the first step where
    they disagree by more than `tol` a 1e-8 comparison says nothing about correctness."""
    n = min(nmax, len(a1), len(a2))
    d = np.maximum(np.abs(a1[:n] - a2[:n]), np.abs(b1[:n] - b2[:n]))
    bad = np.nonzero(d > tol)[0]
    return int(bad[0]) if len(bad) else n


def _solver(cfg, sparse=True, algo=None):
    s = edgpu.Solver(ed_sparse_h=sparse, device=0, **configs.solver_kwargs(cfg))
    if algo is not None:
        s.set_option("hxv_algo", algo)
    return s


def _c4_hloc():
    h = np.zeros((1, 1, 2, 2))
    h[0, 0, 0, 1] = h[0, 0, 1, 0] = 0.3
    h[0, 0, 0, 0], h[0, 0, 1, 1] = 0.1, -0.2
    return h


CASES = [("C1", (4, 4), {}), ("C1", (5, 3), {}), ("C1", (0, 8), {}), ("C1", (8, 1), {}), ("NS6", (3, 2), {}),
         ("NS10", (5, 5), {}), ("NS12", (6, 6), {}), ("NS12", (7, 4), {}),
         ("C4", (5, 5), {}), ("C4", (6, 5), {}), ("C4", (4, 3), {"imphloc": _c4_hloc()})]


@pytest.mark.parametrize("name,sec,over", CASES)
def test_basis_and_csr_bit_exact(name, sec, over):
    cfg, o = make_oracle(name, **over)
    s = _solver(cfg)
    try:
        with o.sector(*sec) as os_:
            s.build_Hv_sector(s.get_sector(*sec))
            assert s.nloc == os_.dim == s.vecDim_Hv_sector(s.get_sector(*sec))
            assert np.array_equal(s.sector_map(0), os_.map_up())
            assert np.array_equal(s.sector_map(1), os_.map_dw())
            for which, ref in ((0, os_.hup()), (1, os_.hdw()), (2, os_.hnd())):
                rp, cols, vals = s.csr(which)
                if which == 2 and cfg["norb"] == 1:
                    assert len(cols) == 0
                    continue
                assert np.array_equal(rp, ref[0]) and np.array_equal(cols, ref[1])
                assert np.array_equal(vals, ref[2])            # +-amplitude: exact
            assert np.array_equal(s.diag(), os_.h0d())         # stored diagonal: same summation order, no FMA
            s.delete_Hv_sector()
    finally:
        s.close()


@pytest.mark.parametrize("name,sec,over", CASES)
@pytest.mark.parametrize("sparse", [True, False])
def test_hxv_matches_oracle(name, sec, over, sparse):
    cfg, o = make_oracle(name, **over)
    s = _solver(cfg, sparse, edgpu.ALGO_GATHER)
    try:
        with o.sector(*sec) as os_:
            v = configs.bench_vector(os_.dim)
            v /= np.linalg.norm(v)
            ref = os_.spmatvec(v)
            s.build_Hv_sector(s.get_sector(*sec))
            hv = s.spHtimesV(v)
            scale = max(np.abs(ref).max(), 1e-300)
            assert np.abs(hv - ref).max() < 1e-13 * scale
            if not sparse:
                assert np.abs(s.diag() - os_.h0d()).max() < 1e-13 * max(1.0, np.abs(os_.h0d()).max())
            hv2 = s.spHtimesV_fortran(v)                       # procedure-pointer compatible symbol
            assert np.array_equal(hv, hv2)
            s.delete_Hv_sector()
    finally:
        s.close()


def test_lifecycle_errors():
    cfg, _ = make_oracle("C1")
    s = _solver(cfg)
    try:
        with pytest.raises(edgpu.EdgpuError):                  # Hstatus = F
            s.nloc = 10
            s.spHtimesV(np.zeros(10))
        isec = s.get_sector(4, 4)
        s.build_Hv_sector(isec)
        with pytest.raises(edgpu.EdgpuError):                  # double build
            s.build_Hv_sector(isec)
        with pytest.raises(edgpu.EdgpuError):                  # Nloc /= dim
            s.spHtimesV(np.zeros(17))
        s.delete_Hv_sector()
        s.build_Hv_sector(s.get_sector(3, 3))                  # rebuild after delete is fine
        assert s.nloc == 56 * 56
        s.delete_Hv_sector()
    finally:
        s.close()


@pytest.mark.parametrize("name,sec", [("C1", (4, 4)), ("NS10", (5, 5)), ("C4", (5, 5))])
@pytest.mark.parametrize("sparse", [True, False])
def test_lanczos_ground_state(name, sec, sparse):
    cfg, o = make_oracle(name)
    s = _solver(cfg, sparse)
    try:
        with o.sector(*sec) as os_:
            v0 = np.ones(os_.dim) / np.sqrt(os_.dim)
            e_ref, vec_ref, a_ref, b_ref = os_.lanc_eigh(v0=v0)
            s.build_Hv_sector(s.get_sector(*sec))
            e0, vec, a, b = s.sp_lanc_eigh(v0)
            assert abs(e0 - e_ref) < 1e-12 * abs(e_ref)
            n = min(50, len(a), len(a_ref))
            assert np.abs(a[:n] - a_ref[:n]).max() < 1e-8
            assert np.abs(b[:n] - b_ref[:n]).max() < 1e-8
            assert abs(np.linalg.norm(vec) - 1) < 1e-12
            ov = abs(vec @ vec_ref)
            assert abs(ov - 1) < 1e-10                          # same eigenvector up to sign
            assert np.linalg.norm(os_.spmatvec(vec) - e0 * vec) < 1e-6
            # zero start vector -> pseudo-random start (no symmetry constraint: may land BELOW the energy
            # reached from the uniform vector, which stays in the totally symmetric subspace)
            e0r, _, _, _ = s.sp_lanc_eigh()
            e_rand, _, _, _ = os_.lanc_eigh()
            assert abs(e0r - e_rand) < 1e-10 * abs(e_rand)
            assert e0r <= e_ref + 1e-10
            s.delete_Hv_sector()
    finally:
        s.close()


def test_c1_cross_check_value():
    cfg, _ = make_oracle("C1")
    s = _solver(cfg)
    try:
        s.build_Hv_sector(s.get_sector(4, 4))
        e0, _, _, _ = s.sp_lanc_eigh(np.ones(4900) / 70.0)
        assert abs(e0 - (-9.361735245469)) < 1e-11
        s.delete_Hv_sector()
    finally:
        s.close()


def test_lanczos_tridiag_coefficients():
    cfg, o = make_oracle("NS10")
    s = _solver(cfg)
    try:
        with o.sector(6, 5) as os_:
            v = configs.bench_vector(os_.dim)
            a_ref, b_ref = os_.lanc_tridiag(v, 60)
            s.build_Hv_sector(s.get_sector(6, 5))
            a, b = s.sp_lanc_tridiag(v, 60)
            assert np.abs(a[:50] - a_ref[:50]).max() < 1e-8
            assert np.abs(b[:50] - b_ref[:50]).max() < 1e-8
            assert b[0] == 0.0
            s.delete_Hv_sector()
    finally:
        s.close()


def test_tridiag_small_sector_exhausts_krylov_space():
    """nlanc = min(jdim, ngfiter); an invariant subspace is hit early -> |b| < threshold exit."""
    cfg, o = make_oracle("NS6")
    s = _solver(cfg)
    try:
        with o.sector(1, 0) as os_:                             # dim 6
            v = np.ones(os_.dim)
            a_ref, b_ref = os_.lanc_tridiag(v, 6)
            s.build_Hv_sector(s.get_sector(1, 0))
            a, b = s.sp_lanc_tridiag(v, 6)
            assert np.abs(a - a_ref).max() < 1e-8 and np.abs(b - b_ref).max() < 1e-8
            s.delete_Hv_sector()
    finally:
        s.close()


@pytest.mark.parametrize("name", ["C1", "C4"])
def test_gf_chains_g_and_sigma(name):
    """Batched GF chains on device vs lanc_build_gf_normal_main of the oracle; G and Sigma to 1e-8."""
    import oracle as O
    cfg, o = make_oracle(name)
    nup, ndw = cfg["nup"], cfg["ndw"]
    with o.sector(nup, ndw) as os_:
        e0, gs, _, _ = os_.lanc_eigh(v0=np.ones(os_.dim) / np.sqrt(os_.dim))
    lmats = 128
    s = _solver(cfg)
    try:
        s.gf_set_state(s.get_sector(nup, ndw), gs, e0)
        chans = []
        for iorb in range(1, cfg["norb"] + 1):
            chans += [(iorb, 1, +1), (iorb, 1, -1)]
        res = s.gf_chains(chans, nlanc_max=200)
        for iorb in range(1, cfg["norb"] + 1):
            ref = o.build_gf_normal(nup, ndw, gs, e0, iorb, lmats=lmats, lreal=32)
            ref_direct = o.build_gf_normal(nup, ndw, gs, e0, iorb, lmats=lmats, lreal=32, mode=1)
            assert np.abs(ref_direct["gmats"] - ref["gmats"]).max() < 1e-9      # the observable is stable
            z = 1j * ref["wm"]
            g = np.zeros(lmats, dtype=complex)
            for k, isign in ((0, 1), (1, -1)):
                r = res[2 * (iorb - 1) + k]
                rc = ref["chains"][k]
                rc2 = ref_direct["chains"][k]
                assert r["nlanc"] == rc["nlanc"]
                assert abs(r["norm2"] - rc["norm2"]) < 1e-12
                ns_ = stable_steps(rc["alanc"], rc["blanc"], rc2["alanc"], rc2["blanc"])
                assert ns_ >= 25, ns_
                assert np.abs(r["alanc"][:ns_] - rc["alanc"][:ns_]).max() < 1e-8
                assert np.abs(r["blanc"][:ns_] - rc["blanc"][:ns_]).max() < 1e-8
                g += edgpu.add_to_lanczos_gf(r["norm2"], e0, r["alanc"], r["blanc"], isign, z)
            assert np.abs(g - ref["gmats"]).max() < 1e-8
            sig, _ = edgpu.sigma_normal(z, g, cfg["xmu"], 0.0, cfg["bath_e"][0, iorb - 1], cfg["bath_v"][0, iorb - 1])
            sig_ref, _ = o.sigma_normal(iorb, 1, z, ref["gmats"])
            assert np.abs(sig - sig_ref).max() < 1e-8
    finally:
        s.close()


@pytest.mark.parametrize("name", ["C1", "C4"])
def test_gf_chains_spin_down_and_device_state(name):
    """Spin-down channels (c / c^+ on the dw word = signed column permutation of the state) and the device-resident
    hand-off of the ground state from sp_lanc_eigh to the chains (no host round trip): norm2, a_n, b_n against the
    oracle's lanc_build_gf_normal_main with ispin = 2 (Nspin = 1 parameters, so G_dw = G_up is an extra check)."""
    cfg, o = make_oracle(name)
    nup, ndw = cfg["nup"], cfg["ndw"]
    s = _solver(cfg)
    try:
        with o.sector(nup, ndw) as os_:
            v0 = np.ones(os_.dim) / np.sqrt(os_.dim)
            e_ref, gs_ref, _, _ = os_.lanc_eigh(v0=v0)
        isec = s.get_sector(nup, ndw)
        s.build_Hv_sector(isec)
        e0, vec, _, _ = s.sp_lanc_eigh(v0)
        s.gf_set_state_from_eigh()                              # eigenvector stays on the device
        s.delete_Hv_sector()
        sign = 1.0 if vec @ gs_ref > 0 else -1.0                 # the chains are invariant under gs -> -gs
        chans = [(io, sp, ar) for io in range(1, cfg["norb"] + 1) for sp in (1, 2) for ar in (1, -1)]
        res = s.gf_chains(chans, nlanc_max=120)
        k = 0
        for io in range(1, cfg["norb"] + 1):
            for sp in (1, 2):
                ref = o.build_gf_normal(nup, ndw, sign * vec, e0, io, ispin=sp, ngfiter=120, lmats=16, lreal=8)
                for p_ in range(2):
                    r, rc = res[k], ref["chains"][p_]
                    k += 1
                    assert r["nlanc"] == rc["nlanc"]
                    assert abs(r["norm2"] - rc["norm2"]) < 1e-12
                    assert np.abs(r["alanc"][:25] - rc["alanc"][:25]).max() < 1e-8
                    assert np.abs(r["blanc"][:25] - rc["blanc"][:25]).max() < 1e-8
        assert abs(e0 - e_ref) < 1e-12 * abs(e_ref)
    finally:
        s.close()


@pytest.mark.parametrize("name", ["C1", "NS6V"])
def test_sector_scan_matches_oracle(name):
    """ed_diag_d's loop over all (Nup, Ndw) sectors on the device (edgpu_diag_sectors): E0 of every sector against dense
    eigh of the oracle's Hmat (small sectors) or the oracle's Lanczos, twin sectors reused, and the global ground state
    left on the device drives the observables without a host round trip."""
    cfg, o = make_oracle(name)
    ns = cfg["nbath"] + 1
    s = _solver(cfg)
    try:
        pairs = [(nu, nd) for nu in range(ns + 1) for nd in range(ns + 1)]
        secs = [s.get_sector(nu, nd) for nu, nd in pairs]
        e0, nl, best = s.diag_sectors(secs, twin=True)
        ref = np.zeros(len(pairs))
        for k, (nu, nd) in enumerate(pairs):
            if nu < nd:
                continue
            with o.sector(nu, nd) as os_:
                if os_.dim <= 800:
                    ref[k] = np.linalg.eigvalsh(os_.hmat())[0]
                else:
                    ref[k] = os_.lanc_eigh(v0=None)[0]
        for k, (nu, nd) in enumerate(pairs):
            if nu < nd:
                ref[k] = ref[pairs.index((nd, nu))]
        assert np.abs(e0 - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())
        assert best == int(np.argmin(e0)) or abs(e0[best] - e0.min()) < 1e-12
        nu, nd = pairs[best]
        with o.sector(nu, nd) as os_:
            _, gs, _, _ = os_.lanc_eigh()                     # pseudo-random start: no symmetry constraint
        obs, oref = s.observables(), o.observables(nu, nd, gs)
        assert np.abs(obs["dens"] - oref["dens"]).max() < 1e-7 and np.abs(obs["docc"] - oref["docc"]).max() < 1e-7
    finally:
        s.close()


def _c4_offdiag_hloc():
    h = np.zeros((1, 1, 2, 2))
    h[0, 0, 0, 1] = h[0, 0, 1, 0] = 0.3
    h[0, 0, 0, 0], h[0, 0, 1, 1] = 0.1, -0.2
    return h


@pytest.mark.parametrize("name,over", [("C1", {}), ("NS10V", {}), ("C4", {}), ("C4", {"imphloc": _c4_offdiag_hloc()})])
def test_observables_match_oracle(name, over):
    """One fused device reduction over the resident ground state (observables.cu) against the oracle's restatement
    of lanc_observables / lanc_local_energy (ED_OBSERVABLES.f90:95-363, 372-600): every output to 1e-12."""
    cfg, o = make_oracle(name, **over)
    nup, ndw = cfg["nup"], cfg["ndw"]
    with o.sector(nup, ndw) as os_:
        e0, gs, _, _ = os_.lanc_eigh(v0=np.ones(os_.dim) / np.sqrt(os_.dim))
    ref = o.observables(nup, ndw, gs, zeta=1.0)
    s = _solver(cfg)
    try:
        s.gf_set_state(s.get_sector(nup, ndw), gs, e0)
        got = s.observables(zeta=1.0)
        for key, val in ref.items():
            assert np.abs(np.asarray(got[key]) - np.asarray(val)).max() < 1e-12, key
        if cfg["xmu"] == 0.0 and not over:                     # half filling with HFMODE: <n> = 1 per orbital
            assert np.abs(got["dens"][:cfg["norb"]] - 1.0).max() < 1e-8
        assert abs(got["prob"].sum() - 1.0) < 1e-10
    finally:
        s.close()


def test_golden_fixtures_on_gpu():
    gold = np.load(os.path.join(ROOT, "tests", "golden", "c1_golden.npz"))
    cfg = configs.config("C1")
    s = _solver(cfg)
    try:
        s.build_Hv_sector(s.get_sector(4, 4))
        v = configs.bench_vector(4900)
        v /= np.linalg.norm(v)
        assert np.abs(s.spHtimesV(v) - gold["hv"]).max() < 1e-14
        assert np.array_equal(s.sector_map(0), gold["map_up"])
        rp, cols, vals = s.csr(0)
        assert np.array_equal(rp, gold["hup_rowptr"]) and np.array_equal(cols, gold["hup_cols"])
        assert np.array_equal(s.diag(), gold["h0d"])
        e0, vec, a, b = s.sp_lanc_eigh(np.ones(4900) / 70.0)
        assert abs(e0 - gold["e0"]) < 1e-12 * abs(gold["e0"])
        s.delete_Hv_sector()
        s.gf_set_state(s.get_sector(4, 4), gold["gs"], float(gold["e0"]))
        res = s.gf_chains([(1, 1, 1), (1, 1, -1)], nlanc_max=200)
        z = 1j * gold["wm"]
        g = sum(edgpu.add_to_lanczos_gf(r["norm2"], float(gold["e0"]), r["alanc"], r["blanc"], sg, z)
                for r, sg in zip(res, (1, -1)))
        assert np.abs(g - gold["gmats"]).max() < 1e-8
    finally:
        s.close()
    gold4 = np.load(os.path.join(ROOT, "tests", "golden", "c4_golden.npz"))
    cfg = configs.config("C4")
    s = _solver(cfg)
    try:
        s.build_Hv_sector(s.get_sector(5, 5))
        v = configs.bench_vector(63504)
        v /= np.linalg.norm(v)
        hv = s.spHtimesV(v)
        assert np.abs(hv - gold4["hv"]).max() < 1e-13 * np.abs(gold4["hv"]).max()
        rp, cols, vals = s.csr(2)
        assert np.array_equal(rp, gold4["hnd_rowptr"]) and np.array_equal(cols, gold4["hnd_cols"])
        assert np.array_equal(vals, gold4["hnd_vals"])
        s.delete_Hv_sector()
    finally:
        s.close()


@pytest.mark.parametrize("name", ["C2", "C3"])
def test_full_size_properties(name):
    """BASELINE full sizes, device-resident, size-independent properties: <x,Hy> = <Hx,y>, stored == direct,
    linearity, and a column-block spot check against the oracle's spMatVec_main loops."""
    import ctypes as C
    cfg = configs.config(name)
    nup, ndw = cfg["nup"], cfg["ndw"]
    outs = {}
    for sparse in (True, False):
        s = _solver(cfg, sparse, edgpu.ALGO_GATHER)
        try:
            s.build_Hv_sector(s.get_sector(nup, ndw))
            n = s.nloc
            dx, dy, dhx, dhy = [s.dev_alloc(8 * n) for _ in range(4)]
            s.dev_fill_bench_vector(dx, n, 0)
            s.dev_fill_bench_vector(dy, n, 12345)
            s.hxv_device(dx, dhx)
            s.hxv_device(dy, dhy)
            s.sync()
            # read back strided samples for the inner products (full vectors for C2)
            x = np.empty(n); y = np.empty(n); hx = np.empty(n); hy = np.empty(n)
            for d, h in ((dx, x), (dy, y), (dhx, hx), (dhy, hy)):
                s.dev_download(d, h)
            lhs, rhs = x @ hy, hx @ y
            assert abs(lhs - rhs) < 1e-10 * max(abs(lhs), np.linalg.norm(x) * np.linalg.norm(hy) * 1e-3)
            outs[sparse] = hx
            for d in (dx, dy, dhx, dhy):
                s.dev_free(d)
            s.delete_Hv_sector()
        finally:
            s.close()
    scale = np.abs(outs[True]).max()
    assert np.abs(outs[True] - outs[False]).max() < 1e-12 * scale
    # spot checks against the oracle's spMatVec_main loops: column blocks spread over the whole i_dw range
    # (ranks 0, 85, 170, 255 of 256 = ~13-50 columns each, both ends included)
    _, o = make_oracle(name)
    for r in (0, 85, 170, 255):
        with o.sector(nup, ndw, r, 256) as blk:
            ref = blk.spmatvec_block(x)
            assert np.abs(outs[True][blk.ishift:blk.ishift + blk.nloc] - ref).max() < 1e-13 * scale


TILED_CASES = [("C1", (4, 4), {}), ("C1", (5, 3), {}), ("C1", (3, 6), {"col_h": 2, "tile_h": 3}),
               ("NS10", (5, 5), {"col_h": 3, "tile_h": 2}), ("NS12", (6, 6), {"col_h": 4, "tile_h": 5}),
               ("NS12", (7, 4), {"col_h": 2, "tile_h": 7, "tile_rows": 8}), ("NS12", (6, 6), {"tile_rows": 8}),
               ("NS6", (3, 3), {"col_h": 5, "tile_h": 6})]


@pytest.mark.parametrize("name,sec,opts", TILED_CASES)
@pytest.mark.parametrize("sparse", [True, False])
def test_tiled_hxv_matches_oracle(name, sec, opts, sparse):
    """Shared-memory staged two-pass kernels, including chunkings that force cross-chunk hops."""
    cfg, o = make_oracle(name)
    s = _solver(cfg, sparse, edgpu.ALGO_TILED)
    for k, v in opts.items():
        s.set_option(k, v)
    try:
        with o.sector(*sec) as os_:
            v = configs.bench_vector(os_.dim)
            v /= np.linalg.norm(v)
            ref = os_.spmatvec(v)
            s.build_Hv_sector(s.get_sector(*sec))
            hv = s.spHtimesV(v)
            assert np.abs(hv - ref).max() < 1e-13 * np.abs(ref).max()
            e_ref, _, a_ref, b_ref = os_.lanc_eigh(v0=np.ones(os_.dim) / np.sqrt(os_.dim))
            e0, _, a, b = s.sp_lanc_eigh(np.ones(os_.dim) / np.sqrt(os_.dim))
            assert abs(e0 - e_ref) < 1e-12 * abs(e_ref)
            s.delete_Hv_sector()
    finally:
        s.close()


# srow_t = t + 1 forces chunks of 2^t low groups (t = 0: every high-bit hop reads L2, none the tile)
FAST_CASES = [("C1", (4, 4), {}), ("C1", (5, 3), {"srow_t": 1}), ("C1", (3, 6), {"srow_lr": 4, "srow_t": 2}),
              ("NS10", (5, 5), {"srow_lr": 4, "srow_t": 3}), ("NS10", (5, 5), {"srow_lr": 5, "srow_t": 2}),
              ("NS12", (6, 6), {"srow_t": 4}), ("NS12", (7, 4), {}), ("NS12", (6, 6), {"no_uniform": 1, "srow_lr": 4}), ("NS12", (6, 6), {"no_uniform": 2}),
              ("NS6", (3, 3), {}), ("NS6", (3, 3), {"srow_lr": 4, "srow_t": 1}), ("NS10V", (5, 5), {"srow_t": 3}),
              ("NS12V", (6, 5), {"srow_lr": 4, "srow_t": 4}), ("NS10V", (7, 2), {}), ("NS10", (9, 1), {}),
              ("NS12", (6, 0), {}), ("NS12", (6, 12), {}), ("NS14", (7, 7), {}), ("NS14V", (7, 6), {"srow_t": 3}),
              # two-CTA cluster column kernel (columns too long for one SM; forced here on small ones)
              ("C1", (4, 4), {"col_cluster": 1}), ("NS12", (6, 6), {"col_cluster": 1}), ("NS12", (7, 4), {"col_cluster": 1, "no_uniform": 1}),
              ("NS10V", (5, 5), {"col_cluster": 1}), ("NS12", (5, 6), {"col_cluster": 1, "srow_lr": 4})]


@pytest.mark.parametrize("name,sec,opts", FAST_CASES)
@pytest.mark.parametrize("sparse", [True, False])
def test_fast_hxv_matches_oracle(name, sec, opts, sparse):
    """TMA-staged whole-column kernel + structured row kernel (hxv_fast.cu), including chunkings that
    force out-of-chunk hops, both low-group widths, level-dependent V_k and edge sectors."""
    cfg, o = make_oracle(name)
    s = _solver(cfg, sparse, edgpu.ALGO_FAST)
    for k, v in opts.items():
        s.set_option(k, v)
    try:
        with o.sector(*sec) as os_:
            v = configs.bench_vector(os_.dim)
            v /= np.linalg.norm(v)
            ref = os_.spmatvec(v)
            s.build_Hv_sector(s.get_sector(*sec))
            hv = s.spHtimesV(v)
            assert np.abs(hv - ref).max() < 1e-13 * max(np.abs(ref).max(), 1e-300)
            if 1 < os_.dim < 2000000:
                e_ref, _, a_ref, b_ref = os_.lanc_eigh(v0=np.ones(os_.dim) / np.sqrt(os_.dim))
                e0, _, a, b = s.sp_lanc_eigh(np.ones(os_.dim) / np.sqrt(os_.dim))
                assert abs(e0 - e_ref) < 1e-12 * abs(e_ref)
            s.delete_Hv_sector()
    finally:
        s.close()


# (config, sector, ranks, srow_lr, srow_t, col_cluster)
SHARD_CASES = [("C1", (4, 4), 2, 0, 0, 0), ("C1", (4, 4), 3, 0, 0, 0), ("C1", (4, 4), 4, 0, 1, 0), ("C1", (5, 3), 4, 0, 2, 0),
               ("NS10", (5, 5), 3, 4, 3, 0), ("NS10", (5, 5), 8, 0, 0, 0), ("NS10", (6, 4), 5, 0, 0, 1), ("NS12", (6, 6), 7, 0, 4, 0),
               ("NS12", (7, 5), 4, 4, 0, 0), ("NS10V", (5, 5), 6, 0, 0, 0), ("NS12V", (6, 5), 8, 4, 3, 1), ("C1", (3, 6), 8, 4, 1, 0),
               ("NS14", (7, 7), 8, 0, 0, 0), ("NS14V", (7, 6), 5, 0, 4, 0)]


@pytest.mark.parametrize("name,sec,nranks,lr,st,cluster", SHARD_CASES)
@pytest.mark.parametrize("sparse", [True, False])
def test_sharded_fast_path_emulated_on_one_gpu(name, sec, nranks, lr, st, cluster, sparse):
    """The kernels and plans of the multi-GPU fast path (whole / cut low groups, group records, the halo copy
    kernel, the column pass's source lists for remote and boundary hops) with the ranks emulated by contexts on ONE
    device (edgpu_selftest_sharded_hxv): any rank count 1..8, including splits where DimDw is not a multiple of P
    and ranks that own no whole low group."""
    import ctypes as C
    cfg, o = make_oracle(name)
    s = _solver(cfg, sparse)
    try:
        with o.sector(*sec) as os_:
            v = configs.bench_vector(os_.dim)
            v /= np.linalg.norm(v)
            ref = os_.spmatvec(v)
            out = np.zeros(os_.dim)
            L = edgpu.selftest_lib()
            L.edgpu_selftest_sharded_hxv.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64,
                                                     C.c_void_p, C.c_void_p]
            rc = L.edgpu_selftest_sharded_hxv(C.byref(s._keep[0]), sec[0], sec[1], nranks, lr, st, cluster,
                                              v.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
            assert rc == 0, edgpu.lib().edgpu_last_error().decode()
            assert np.abs(out - ref).max() < 1e-13 * np.abs(ref).max()
    finally:
        s.close()


def test_fast_unsupported_is_loud():
    """Odd DimUp (TMA bulk copies need 16-byte aligned columns): the explicit fast algorithm refuses,
    AUTO falls back to the generic kernels."""
    cfg, o = make_oracle("NS7")
    s = _solver(cfg, True, edgpu.ALGO_FAST)
    try:
        s.build_Hv_sector(s.get_sector(3, 3))          # DimUp = C(7,3) = 35
        v = np.ones(35 * 35)
        with pytest.raises(Exception):
            s.spHtimesV(v)
        s.set_option("hxv_algo", edgpu.ALGO_AUTO)
        with o.sector(3, 3) as os_:
            assert np.abs(s.spHtimesV(v) - os_.spmatvec(v)).max() < 1e-12
        s.delete_Hv_sector()
    finally:
        s.close()


@pytest.mark.parametrize("name", ["C2", "C3"])
def test_fast_equals_gather_full_size(name):
    cfg = configs.config(name)
    out = {}
    for algo in (edgpu.ALGO_GATHER, edgpu.ALGO_FAST):
        for sparse in ((False, True) if algo == edgpu.ALGO_FAST else (False,)):
            s = _solver(cfg, sparse, algo)
            try:
                s.build_Hv_sector(s.get_sector(cfg["nup"], cfg["ndw"]))
                n = s.nloc
                dx, dy = s.dev_alloc(8 * n), s.dev_alloc(8 * n)
                s.dev_fill_bench_vector(dx, n, 0)
                s.hxv_device(dx, dy)
                s.sync()
                out[(algo, sparse)] = np.empty(n)
                s.dev_download(dy, out[(algo, sparse)])
                s.dev_free(dx)
                s.dev_free(dy)
                s.delete_Hv_sector()
            finally:
                s.close()
    ref = out[(edgpu.ALGO_GATHER, False)]
    scale = np.abs(ref).max()
    assert np.abs(ref - out[(edgpu.ALGO_FAST, False)]).max() < 1e-12 * scale
    assert np.abs(ref - out[(edgpu.ALGO_FAST, True)]).max() < 1e-12 * scale


def test_c5_ns18_on_one_gpu():
    """BASELINE configs[4]: Ns=18, sector 9:9, dim 2 363 904 400 (19 GB per vector) on ONE B200.  The fast path
    (aligned-chunk row kernel + 2-CTA cluster column kernel) equals the one-pass gather kernel on the whole vector
    (compared on the device), and both ends plus the middle of the i_dw range agree with the oracle's spMatVec_main
    loops (oracle/, sparse-column form: the full vector does not have to exist on the host)."""
    import ctypes as C
    cfg = configs.config("C5")
    _, o = make_oracle("C5")
    s = _solver(cfg, False, edgpu.ALGO_FAST)
    try:
        s.build_Hv_sector(s.get_sector(cfg["nup"], cfg["ndw"]))
        n = s.nloc
        assert n == 2363904400
        dx, dyf, dyg = s.dev_alloc(8 * n), s.dev_alloc(8 * n), s.dev_alloc(8 * n)
        s.dev_fill_bench_vector(dx, n, 0)
        s.hxv_device(dx, dyf)
        s.set_option("hxv_algo", edgpu.ALGO_GATHER)
        s.hxv_device(dx, dyg)
        s.sync()
        T = edgpu.selftest_lib()
        T.edgpu_selftest_dev_maxabsdiff.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        md, ma = C.c_double(0), C.c_double(0)
        assert T.edgpu_selftest_dev_maxabsdiff(dyf, dyg, n, C.byref(md), C.byref(ma)) == 0
        assert ma.value > 1.0 and md.value < 1e-12 * ma.value, (md.value, ma.value)
        for j in (0, 1, s.dimdw // 3, s.dimdw // 2, s.dimdw - 2, s.dimdw - 1):
            with o.sector(cfg["nup"], cfg["ndw"], j, s.dimdw) as blk:
                ci = blk.block_columns()
                xc = np.stack([configs.bench_vector(s.dimup, int(q) * s.dimup) for q in ci])
                ref = blk.spmatvec_block_cols(ci, xc)
            got = np.empty(s.dimup)
            s.dev_download_slice(dyf, j * s.dimup, got)
            assert np.abs(got - ref).max() < 1e-13 * ma.value, j
        for d in (dx, dyf, dyg):
            s.dev_free(d)
        s.delete_Hv_sector()
    finally:
        s.close()


@pytest.mark.parametrize("name", ["C2", "C3"])
def test_tiled_equals_gather_full_size(name):
    cfg = configs.config(name)
    out = {}
    for algo in (edgpu.ALGO_GATHER, edgpu.ALGO_TILED):
        s = _solver(cfg, False, algo)
        try:
            s.build_Hv_sector(s.get_sector(cfg["nup"], cfg["ndw"]))
            n = s.nloc
            dx, dy = s.dev_alloc(8 * n), s.dev_alloc(8 * n)
            s.dev_fill_bench_vector(dx, n, 0)
            s.hxv_device(dx, dy)
            s.sync()
            out[algo] = np.empty(n)
            s.dev_download(dy, out[algo])
            s.dev_free(dx)
            s.dev_free(dy)
            s.delete_Hv_sector()
        finally:
            s.close()
    scale = np.abs(out[edgpu.ALGO_GATHER]).max()
    assert np.abs(out[edgpu.ALGO_GATHER] - out[edgpu.ALGO_TILED]).max() < 1e-12 * scale


@pytest.mark.parametrize("name", ["C1", "C4"])
def test_chi_chains_match_oracle(name):
    """Spin and density susceptibility chains (edgpu_chi_chains: lanc_ed_build_spinChi_* / lanc_ed_build_densChi_*,
    ED_GF_CHISPIN.f90:114-415, ED_GF_CHIDENS.f90:111-426) from the device-resident state, every channel kind (one
    orbital, total, mixed), batched: norm2 to 1e-12, a_n / b_n to 1e-8 over the stable prefix, and
    add_to_lanczos_chi's chi(i nu), chi(tau), chi(w) to 1e-8 against the oracle."""
    cfg, o = make_oracle(name)
    nup, ndw = cfg["nup"], cfg["ndw"]
    s = _solver(cfg)
    try:
        with o.sector(nup, ndw) as os_:
            v0 = np.ones(os_.dim) / np.sqrt(os_.dim)
            e_ref, gs_ref, _, _ = os_.lanc_eigh(v0=v0)
        isec = s.get_sector(nup, ndw)
        s.gf_set_state(isec, gs_ref, e_ref)
        chans = [(io, io) for io in range(1, cfg["norb"] + 1)] + [(0, 0)]
        if cfg["norb"] > 1:
            chans.append((1, 2))
        beta = 50.0
        vm = np.pi / beta * 2 * np.arange(9)
        tau = np.linspace(0.0, beta, 11)
        vr = np.linspace(-3, 3, 13)
        for kind in (0, 1):
            res = s.chi_chains(kind, chans, nlanc_max=100)
            for (io, jo), r in zip(chans, res):
                ref = o.chi_chain(nup, ndw, gs_ref, kind, io, jo, ngfiter=100)
                assert r["nlanc"] == ref["nlanc"]
                assert abs(r["norm2"] - ref["norm2"]) < 1e-12
                # the density operator overlaps the ground state itself, so its chain breaks down early: compare the
                # coefficients while the oracle's b_n stay away from zero
                m = 1
                while m < min(20, ref["nlanc"]) and abs(ref["blanc"][m]) > 1e-3:
                    m += 1
                assert np.abs(r["alanc"][:m] - ref["alanc"][:m]).max() < 1e-8
                assert np.abs(r["blanc"][:m] - ref["blanc"][:m]).max() < 1e-8
                got = edgpu.add_to_lanczos_chi(r["norm2"], e_ref, beta, r["alanc"], r["blanc"], vm, tau, vr, 0.01)
                want = O.add_to_lanczos_chi(ref["norm2"], e_ref, beta, ref["alanc"], ref["blanc"], vm, tau, vr, 0.01)
                for g_, w_ in zip(got[:2], want[:2]):              # chi(i nu), chi(tau): the stable outputs
                    assert np.abs(g_ - w_).max() < 1e-8
                # chi(w + i eps) resolves single poles, also the unconverged ones at the far end of the Krylov space, which
                # differ between any two correct evaluations: its formula is checked on IDENTICAL coefficients instead --
                # the product's host formula on the ORACLE's coefficients: the two restatements of add_to_lanczos_*Chi agree
                same = edgpu.add_to_lanczos_chi(ref["norm2"], e_ref, beta, ref["alanc"], ref["blanc"], vm, tau, vr, 0.01)
                for g_, w_ in zip(same, want):
                    assert np.abs(g_ - w_).max() < 1e-10
    finally:
        s.close()


ORBS_CASES = [("ORB2", [1, 2], [2, 1], True), ("ORB2", [2, 1], [1, 1], False), ("ORB2", [0, 3], [3, 0], True),
              ("ORB2B", [3, 2], [2, 3], True), ("ORB2B", [2, 2], [3, 3], False), ("ORB3", [2, 1, 2], [1, 2, 1], True)]


@pytest.mark.parametrize("name,nups,ndws,sparse", ORBS_CASES)
def test_orbs_operator_matches_oracle(name, nups, ndws, sparse):
    """ed_total_ud = F (one (Nup, Ndw) pair per orbital): sector number, the 2*Norb word maps and the CSR of every
    spH0ups(iorb) / spH0dws(iorb) bit-exact (structure AND values), the stored diagonal bit-exact, H*v to 1e-13 in both
    the stored and the recomputed-diagonal form, E0 to 1e-12 and the Lanczos coefficients to 1e-8 over the stable prefix
    -- against the oracle's ed_buildh_orbs / spMatVec_orbs restatement (ED_HAMILTONIAN_SPARSE_HxV.f90:206-370, 487-564)."""
    cfg, o = make_oracle(name)
    s = edgpu.Solver(ed_sparse_h=sparse, ed_total_ud=False, **configs.solver_kwargs(cfg))
    try:
        with o.sector_orbs(nups, ndws) as so:
            isec = s.get_sector_orbs(nups, ndws)
            assert isec == so.isector and s.get_qn_orbs(isec) == (list(nups), list(ndws))
            assert s.vecDim_Hv_sector(isec) == so.dim
            s.build_Hv_sector(isec)
            dims, dim = s.orbs_dims()
            assert dims == so.dims and dim == so.dim == s.nloc
            for f in range(2 * cfg["norb"]):
                m, rp, cols, vals = s.orbs_factor(f)
                om, orp, ocols, ovals = so.factor(f)
                assert np.array_equal(m, om) and np.array_equal(rp, orp) and np.array_equal(cols, ocols) and np.array_equal(vals, ovals)
            if sparse:
                assert np.array_equal(s.orbs_diag(), so.h0d())
            v = configs.bench_vector(so.dim)
            v /= np.linalg.norm(v)
            ref = so.spmatvec(v)
            hv = s.spHtimesV(v)
            assert np.abs(hv - ref).max() <= 1e-13 * np.abs(ref).max()
            v0 = np.ones(so.dim) / np.sqrt(so.dim)
            e_ref, vec_ref, a_ref, b_ref = so.lanc_eigh(v0=v0)
            e0, vec, a, b = s.sp_lanc_eigh(v0)
            assert abs(e0 - e_ref) <= 1e-12 * max(1.0, abs(e_ref))
            n = min(15, len(a), len(a_ref))
            assert np.abs(a[:n] - a_ref[:n]).max() < 1e-8 and np.abs(b[:n] - b_ref[:n]).max() < 1e-8
            assert abs(abs(vec @ vec_ref) - 1) < 1e-9
            at, bt = s.sp_lanc_tridiag(v, min(30, so.dim))
            ao, bo = so.lanc_tridiag(v, min(30, so.dim))
            assert np.abs(at[:n] - ao[:n]).max() < 1e-8 and np.abs(bt[:n] - bo[:n]).max() < 1e-8
            s.delete_Hv_sector()
    finally:
        s.close()


def test_orbs_sector_scan_finds_the_total_ground_state():
    """ed_diag_d over every orbital-resolved sector with the totals of the half-filled sector (edgpu_diag_sectors with
    ed_total_ud = F): the lowest E0 equals the ground-state energy of the ed_total_ud = T sector computed by the main
    path -- the two operator families agree on the physics."""
    cfg, o = make_oracle("ORB2")
    nt, nso = cfg["nup"], cfg["nbath"] + 1
    with o.sector(nt, nt) as st:
        e_tot = np.linalg.eigvalsh(st.hmat())[0]
    s = edgpu.Solver(ed_total_ud=False, **configs.solver_kwargs(cfg))
    try:
        secs = [s.get_sector_orbs([nu1, nt - nu1], [nd1, nt - nd1]) for nu1 in range(nso + 1) for nd1 in range(nso + 1)
                if 0 <= nt - nu1 <= nso and 0 <= nt - nd1 <= nso]
        e0, nl, best = s.diag_sectors(secs, threshold=1e-14)
        assert abs(e0.min() - e_tot) < 1e-10 and e0[best] == e0.min()
        with pytest.raises(edgpu.EdgpuError):
            s.gf_chains([(1, 1, 1)])                             # chains of an orbital-resolved state: refused, not faked
    finally:
        s.close()


@pytest.mark.parametrize("name,sec,nph,sparse", [("NS6", (3, 3), 3, True), ("C1", (4, 4), 2, False), ("NS10V", (5, 4), 4, True),
                                                 ("C4", (5, 5), 1, True)])
def test_phonon_operator_matches_oracle(name, sec, nph, sparse):
    """DimPh = Nph + 1 > 1 (one local phonon mode, stored/H_ph.f90 + H_e_ph.f90): the sector vector is v(i_el, iph) with
    the phonon index slowest; the electronic H*v runs on every slab (the fast two-pass path where it applies), the
    phonon energy and the electron-phonon coupling ride on top.  H*v to 1e-13, E0 to 1e-12 and the Lanczos coefficients
    to 1e-8 against the oracle's spMatVec_main with phonons (ED_HAMILTONIAN_SPARSE_HxV.f90:391-485)."""
    cfg, o = make_oracle(name)
    g = tuple(0.3 + 0.2 * k for k in range(cfg["norb"]))
    w0 = 0.5
    s = edgpu.Solver(ed_sparse_h=sparse, nph=nph, g_ph=g, w0_ph=w0, **configs.solver_kwargs(cfg))
    try:
        with o.sector(*sec) as so:
            isec = s.get_sector(*sec)
            n = so.dim * (nph + 1)
            assert s.vecDim_Hv_sector(isec) == n
            s.build_Hv_sector(isec)
            assert s.nloc == n
            v = configs.bench_vector(n)
            v /= np.linalg.norm(v)
            ref = so.spmatvec_ph(v, nph, g, w0)
            hv = s.spHtimesV(v)
            assert np.abs(hv - ref).max() <= 1e-13 * np.abs(ref).max()
            v0 = np.ones(n) / np.sqrt(n)
            e_ref, vec_ref, a_ref, b_ref = so.lanc_eigh_ph(nph, g, w0, v0)
            e0, vec, a, b = s.sp_lanc_eigh(v0)
            assert abs(e0 - e_ref) <= 1e-12 * abs(e_ref)
            m = min(20, len(a), len(a_ref))
            assert np.abs(a[:m] - a_ref[:m]).max() < 1e-8 and np.abs(b[:m] - b_ref[:m]).max() < 1e-8
            assert abs(abs(vec @ vec_ref) - 1) < 1e-9
            s.delete_Hv_sector()
            with pytest.raises(edgpu.EdgpuError):
                s.gf_set_state(isec, vec_ref, e_ref)                 # chains of an electron-phonon state: refused
    finally:
        s.close()


def _bath_models():
    """Small two-orbital models with a hybrid and a replica bath (interaction, spin-exchange and pair-hopping on)."""
    common = dict(norb=2, nspin=1, uloc=(2.0, 1.5), ust=0.8, jh=0.3, jx=0.3, jp=0.3, xmu=0.2, hfmode=True)
    hb = np.zeros((1, 1, 2, 2, 3))
    for kp in range(3):
        hb[0, 0, :, :, kp] = [[-0.8 + 0.7 * kp, 0.25 - 0.1 * kp], [0.25 - 0.1 * kp, 0.3 + 0.2 * kp]]
    replica = dict(nbath=3, bath_type=2, bath_v=np.array([[0.4, 0.7, 0.55]]), bath_h=hb, **common)
    hybrid = dict(nbath=4, bath_type=1, bath_e=np.array([[[-1.0, -0.3, 0.2, 0.9]]]),
                  bath_v=np.array([[[0.5, 0.1], [0.3, 0.4], [0.2, 0.6], [0.45, 0.35]]]).transpose(0, 2, 1), **common)
    return {"replica": (replica, (4, 4)), "hybrid": (hybrid, (3, 3))}


@pytest.mark.parametrize("kind", ["replica", "hybrid"])
@pytest.mark.parametrize("sparse", [True, False])
def test_hybrid_and_replica_baths_match_oracle(kind, sparse):
    """bath_type hybrid / replica (getBathStride ED_SETUP.f90:358-375; replica inter-orbital bath hopping stored/H_up.f90:26-50;
    bath_diag, diag_hybr of ed_buildh_main ED_HAMILTONIAN_SPARSE_HxV.f90:46-76): basis maps and CSR factors bit-exact, stored
    diagonal bit-exact, H*v 1e-13 (stored and direct), E0 1e-12, Lanczos coefficients 1e-8."""
    kw, sec = _bath_models()[kind]
    o = O.Oracle(**kw)
    s = edgpu.Solver(ed_sparse_h=sparse, **kw)
    try:
        with o.sector(*sec) as so:
            isec = s.get_sector(*sec)
            s.build_Hv_sector(isec)
            assert s.nloc == so.dim
            assert np.array_equal(s.sector_map(0), so.map_up()) and np.array_equal(s.sector_map(1), so.map_dw())
            for which, ref in ((0, so.hup()), (1, so.hdw()), (2, so.hnd())):
                rp, cols, vals = s.csr(which)
                assert np.array_equal(rp, ref[0]) and np.array_equal(cols, ref[1]) and np.array_equal(vals, ref[2])
            if sparse:
                assert np.array_equal(s.diag(), so.h0d())
            v = configs.bench_vector(so.dim)
            v /= np.linalg.norm(v)
            ref = so.spmatvec(v)
            hv = s.spHtimesV(v)
            assert np.abs(hv - ref).max() <= 1e-13 * np.abs(ref).max()
            v0 = np.ones(so.dim) / np.sqrt(so.dim)
            e_ref, vec_ref, a_ref, b_ref = so.lanc_eigh(v0=v0)
            e0, vec, a, b = s.sp_lanc_eigh(v0)
            assert abs(e0 - e_ref) <= 1e-12 * abs(e_ref)
            m = min(20, len(a), len(a_ref))
            assert np.abs(a[:m] - a_ref[:m]).max() < 1e-8 and np.abs(b[:m] - b_ref[:m]).max() < 1e-8
            s.delete_Hv_sector()
    finally:
        s.close()


def test_noninteracting_gf_is_the_hybridisation_formula_on_gpu():
    """No oracle in the loop: at U = 0 the engine's own chain -- sector scan on the device (edgpu_diag_sectors), the
    eigenvector handed to the chains without leaving HBM, c / c^+ on the device, batched GF Lanczos, add_to_lanczos_gf --
    must reproduce G(iw) = 1 / (iw + mu - sum_k V_k^2 / (iw - e_k)) to 1e-6, i.e. Sigma = 0 (the closed form behind
    build_sigma_normal, ED_GF_NORMAL.f90:935-1002, ED_BATH_FUNCTIONS.f90:163-195)."""
    cfg = configs.config("NS8V")
    cfg["uloc"] = (0.0,)
    cfg["hfmode"] = False
    kw = configs.solver_kwargs(cfg)
    s = edgpu.Solver(**kw)
    try:
        ns = cfg["nbath"] + 1
        secs = [s.get_sector(nu, nd) for nu in range(ns + 1) for nd in range(ns + 1)]
        e0s, nl, best = s.diag_sectors(secs, threshold=1e-16, twin=False)
        e0 = e0s[best]
        # the same number from the one-body spectrum: fill the negative levels of each spin (the chemical potential acts
        # on the impurity only, stored/H_local.f90:13-18; the bath levels are bare)
        h1 = np.zeros((ns, ns))
        e = np.asarray(cfg["bath_e"]).reshape(-1)
        v = np.asarray(cfg["bath_v"]).reshape(-1)
        h1[0, 0] = -cfg["xmu"]
        for k in range(ns - 1):
            h1[k + 1, k + 1] = e[k]
            h1[0, k + 1] = h1[k + 1, 0] = v[k]
        w1 = np.linalg.eigvalsh(h1)
        assert abs(e0 - 2 * w1[w1 < 0].sum()) < 1e-10
        res = s.gf_chains([(1, 1, +1), (1, 1, -1)], nlanc_max=200)
        beta, lmats = 50.0, 64
        z = 1j * np.pi / beta * (2 * np.arange(1, lmats + 1) - 1)
        g = np.zeros(lmats, dtype=complex)
        for r, isign in zip(res, (1, -1)):
            g += edgpu.add_to_lanczos_gf(r["norm2"], e0, r["alanc"], r["blanc"], isign, z)
        delta = (v[None, :] ** 2 / (z[:, None] - e[None, :])).sum(axis=1)
        # the state is the Lanczos eigenvector of the scan (energy converged to 1e-16, vector to its square root), and
        # G is first order in the vector error
        err = np.abs(g - 1.0 / (z + cfg["xmu"] - delta)).max()
        assert err < 1e-6, err
        sig, _ = edgpu.sigma_normal(z, g, cfg["xmu"], 0.0, e, v)
        assert np.abs(sig).max() < 1e-5, np.abs(sig).max()
        assert abs(res[0]["norm2"] + res[1]["norm2"] - 1.0) < 1e-12
    finally:
        s.close()
