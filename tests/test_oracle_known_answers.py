"""Known-answer tests that pin the CPU oracle (the reference ships no golden
vectors, so these stand in for them).  No GPU needed."""
import numpy as np
import pytest

import oracle as O
from edgpu import configs
from conftest import make_oracle


def test_init_dmft_bath_matches_reference_formula():
    # ED_BATH/dmft_aux.f90:102-133 with HWBAND=2: Nbath=7 -> -2,-4/3,-2/3,0,2/3,4/3,2 ; V = 1/sqrt(7)
    e, v = O.init_dmft_bath(1, 7)
    assert np.allclose(e.ravel(), [-2, -4 / 3, -2 / 3, 0, 2 / 3, 4 / 3, 2], atol=1e-15)
    assert np.allclose(v, 1 / np.sqrt(7))
    e, v = O.init_dmft_bath(2, 4)
    assert np.allclose(e[0, 0], [-2, -0.1, 0.1, 2]) and np.allclose(v, 0.5)
    e2, v2 = configs.init_dmft_bath(2, 4)          # the product's host-side restatement agrees
    assert (e2 == e).all() and (v2 == v).all()
    for nb in (3, 5, 6, 8, 13, 15, 17):
        a, b = O.init_dmft_bath(1, nb)
        c, d = configs.init_dmft_bath(1, nb)
        assert (a == c).all() and (b == d).all()


def test_sector_map_binary_search_binomial():
    L = O.lib()
    for ns, n in [(8, 4), (10, 3), (12, 6), (14, 7)]:
        m = O.build_sector_map(ns, n)
        assert len(m) == L.orc_binomial(ns, n)
        assert (np.diff(m) > 0).all()
        assert all(bin(int(x)).count("1") == n for x in m[:: max(1, len(m) // 200)])
        p = m.ctypes.data_as(O.c_i32p)
        for k in (0, 1, len(m) // 2, len(m) - 1):
            assert L.orc_binary_search(p, len(m), int(m[k])) == k + 1        # 1-based
        assert L.orc_binary_search(p, len(m), int(m[-1]) + 1) == 0           # absent


def test_c_cdg_signs():
    L = O.lib()
    import ctypes as C
    out, sg = C.c_int32(0), C.c_double(0)
    # |0b1011>: destroy site 4 (bit 3): two occupied below -> sign +1 ... site 2 (bit1): one below -> -1
    assert L.orc_c(4, 0b1011, C.byref(out), C.byref(sg)) == 0 and out.value == 0b0011 and sg.value == 1.0
    assert L.orc_c(2, 0b1011, C.byref(out), C.byref(sg)) == 0 and out.value == 0b1001 and sg.value == -1.0
    assert L.orc_c(3, 0b1011, C.byref(out), C.byref(sg)) == 1                 # empty site: error
    assert L.orc_cdg(3, 0b1011, C.byref(out), C.byref(sg)) == 0 and out.value == 0b1111 and sg.value == 1.0
    assert L.orc_cdg(1, 0b1011, C.byref(out), C.byref(sg)) == 1               # occupied: error


@pytest.mark.parametrize("nbath,nup,ndw", [(3, 2, 2), (5, 3, 3), (7, 4, 4), (5, 2, 4)])
def test_u0_free_fermions(nbath, nup, ndw):
    """U=0: E0(sector) = sum of the lowest Nup + lowest Ndw levels of the one-body matrix."""
    cfg = configs.config("NS%d" % (nbath + 1))
    cfg.update(uloc=(0.0,), hfmode=False)
    o = O.Oracle(**configs.solver_kwargs(cfg))
    ns = nbath + 1
    h1 = np.zeros((ns, ns))
    e, v = cfg["bath_e"].ravel(), cfg["bath_v"].ravel()
    h1[1:, 1:] = np.diag(e)
    h1[0, 1:] = v
    h1[1:, 0] = v
    w1 = np.linalg.eigvalsh(h1)
    with o.sector(nup, ndw) as s:
        w = np.linalg.eigvalsh(s.hmat())
    assert abs(w[0] - (w1[:nup].sum() + w1[:ndw].sum())) < 1e-12


def test_atomic_limit_is_diagonal():
    cfg = configs.config("C4")
    cfg["bath_v"] = np.zeros_like(cfg["bath_v"])
    cfg.update(jx=0.0, jp=0.0)
    o = O.Oracle(**configs.solver_kwargs(cfg))
    with o.sector(5, 5) as s:
        rp, cols, vals = s.hup()
        assert rp[-1] == 0
        v = configs.bench_vector(s.dim)
        assert np.array_equal(s.spmatvec(v), 0.0 + s.h0d() * v)


@pytest.mark.parametrize("name,sec", [("C1", (4, 4)), ("C1", (5, 3)), ("NS6", (3, 2)), ("C4", None)])
def test_stored_direct_sharded_dense_agree(name, sec):
    cfg, o = make_oracle(name)
    if name == "C4":
        h = np.zeros((1, 1, 2, 2))
        h[0, 0, 0, 1] = h[0, 0, 1, 0] = 0.3
        h[0, 0, 0, 0], h[0, 0, 1, 1] = 0.1, -0.2
        cfg, o = make_oracle(name, imphloc=h)
        sec = (4, 3)                                  # dim 210*120 = 25200: dense is too big, skip dense
    nup, ndw = sec
    with o.sector(nup, ndw) as s, o.sector(nup, ndw, sparse_h=False) as sd:
        v = configs.bench_vector(s.dim)
        v /= np.linalg.norm(v)
        hv = s.spmatvec(v)
        scale = np.abs(hv).max()
        assert np.abs(sd.directmatvec(v) - hv).max() < 1e-13 * scale
        if s.dim <= 5000:
            H = s.hmat()
            assert np.abs(H - H.T).max() == 0.0
            assert np.abs(H @ v - hv).max() < 1e-13 * scale
        for P in (1, 2, 3, 5, 8):
            if P > s.dimdw:
                continue
            assert np.abs(o.spmatvec_mpi(nup, ndw, P, v, nthreads=2) - hv).max() < 1e-13 * scale
            assert np.abs(o.directmatvec_mpi(nup, ndw, P, v) - hv).max() < 1e-13 * scale
        # <x,Hy> = <Hx,y>
        y = np.cos(0.11 * np.arange(s.dim))
        assert abs(v @ s.spmatvec(y) - hv @ y) < 1e-11 * scale


def test_shard_geometry_matches_reference_rule():
    cfg, o = make_oracle("C1")
    tot = 0
    for r in range(8):
        with o.sector(4, 4, r, 8) as s:                # DimDw=70 = 8*8+6
            assert s.qdw == (9 if r < 6 else 8)
            assert s.ishift == tot and s.nloc == o.vecdim(4, 4, r, 8)
            tot += s.nloc
    assert tot == 4900


def test_vector_transpose_layout():
    """Appendix B of SURVEY.md: after vector_transpose_MPI rank d holds vt(g, r) = V(rowstart_d + r, g)."""
    import ctypes as C
    nrow, ncol, P = 13, 11, 4
    V = np.arange(nrow * ncol, dtype=np.float64).reshape((nrow, ncol), order="F")
    qc = [ncol // P + (1 if r < ncol % P else 0) for r in range(P)]
    qr = [nrow // P + (1 if r < nrow % P else 0) for r in range(P)]
    co = np.cumsum([0] + qc)
    ro = np.cumsum([0] + qr)
    a = [np.asfortranarray(V[:, co[r]:co[r + 1]]).ravel(order="F").copy() for r in range(P)]
    b = [np.zeros(ncol * qr[r]) for r in range(P)]
    pa = (O.c_dp * P)(*[x.ctypes.data_as(O.c_dp) for x in a])
    pb = (O.c_dp * P)(*[x.ctypes.data_as(O.c_dp) for x in b])
    O.lib().orc_vector_transpose_all(P, nrow, ncol, pa, pb)
    for d in range(P):
        vt = b[d].reshape((ncol, qr[d]), order="F")
        assert np.array_equal(vt, V[ro[d]:ro[d + 1], :].T)


def test_tql2_vs_lapack():
    rng = np.random.default_rng(1)
    for n in (1, 2, 7, 50, 200):
        d = rng.normal(size=n)
        e = np.zeros(n)
        e[1:] = rng.normal(size=n - 1)
        w, z = O.tql2(d, e)
        T = np.diag(d) + np.diag(e[1:], 1) + np.diag(e[1:], -1)
        assert np.allclose(w, np.linalg.eigvalsh(T), atol=1e-12)
        assert np.allclose(T @ z, z * w, atol=1e-11)


def test_lanczos_e0_vs_dense_c1():
    """Cross-check value of BASELINE.md section 6: E0 = -9.361735245469, E1 = -8.950104113137."""
    cfg, o = make_oracle("C1")
    with o.sector(4, 4) as s:
        w = np.linalg.eigvalsh(s.hmat())
        assert abs(w[0] - (-9.361735245469)) < 1e-11 and abs(w[1] - (-8.950104113137)) < 1e-11
        e0, vec, a, b = s.lanc_eigh(v0=np.ones(s.dim) / np.sqrt(s.dim))
        assert abs(e0 - w[0]) < 1e-12 * abs(w[0])
        assert abs(np.linalg.norm(vec) - 1) < 1e-12
        assert np.linalg.norm(s.spmatvec(vec) - e0 * vec) < 1e-6
        e0d, _, _, _ = s.lanc_eigh(v0=np.ones(s.dim) / np.sqrt(s.dim), mode=1)
        assert abs(e0d - e0) < 1e-12
        e0r, _, _, _ = s.lanc_eigh()                  # zero start vector -> pseudo-random
        assert abs(e0r - w[0]) < 1e-12 * abs(w[0])


def test_lanczos_tridiag_matches_dense_krylov():
    cfg, o = make_oracle("NS6")
    with o.sector(3, 3) as s:
        H = s.hmat()
        v = configs.bench_vector(s.dim)
        a, b = s.lanc_tridiag(v, 30)
        # explicit Gram-Schmidt Krylov reference
        q = v / np.linalg.norm(v)
        qp = np.zeros_like(q)
        beta = 0.0
        for k in range(12):
            w = H @ q - beta * qp
            alpha = q @ w
            w -= alpha * q
            assert abs(alpha - a[k]) < 1e-10
            beta = np.linalg.norm(w)
            assert abs(beta - b[k + 1]) < 1e-10
            qp, q = q, w / beta


def test_gf_sum_rules_and_half_filling_symmetry():
    cfg, o = make_oracle("C1")
    with o.sector(4, 4) as s:
        e0, gs, _, _ = s.lanc_eigh(v0=np.ones(s.dim) / np.sqrt(s.dim))
    r = o.build_gf_normal(4, 4, gs, e0, 1, lmats=256, lreal=64)
    n_add, n_rem = r["chains"][0]["norm2"], r["chains"][1]["norm2"]
    assert abs(n_add + n_rem - 1.0) < 1e-12            # {c, c^+} = 1
    assert abs(n_rem - 0.5) < 1e-6                     # <n_up> = 1/2 at half filling
    assert np.abs(r["gmats"].real).max() < 1e-6        # particle-hole symmetry
    # high-frequency tail G(iw) -> 1/(iw)
    wm = r["wm"]
    g_big = O.add_to_lanczos_gf(n_add, e0, r["chains"][0]["alanc"], r["chains"][0]["blanc"], 1,
                                np.array([1e6]), np.array([0.0]), 0.01)[0]
    g_big2 = O.add_to_lanczos_gf(n_rem, e0, r["chains"][1]["alanc"], r["chains"][1]["blanc"], -1,
                                 np.array([1e6]), np.array([0.0]), 0.01)[0]
    assert abs((g_big + g_big2)[0] * 1j * 1e6 - 1.0) < 1e-6
    sig, invg0 = o.sigma_normal(1, 1, 1j * wm, r["gmats"])
    assert np.abs(sig.real).max() < 1e-4               # HFMODE: Hartree shift absorbed (GS known to ~1e-8)
    assert (sig.imag < 1e-12).all()


def test_gf_lanczos_vs_full_spectral_sum():
    """ed_diag_type=full analogue (ED_GF_NORMAL.f90:672-790): G(iw) from all eigenpairs of the N+1 / N-1 sectors."""
    cfg, o = make_oracle("NS6")
    nup = ndw = 3
    with o.sector(nup, ndw) as s:
        w, U = np.linalg.eigh(s.hmat())
    e0, gs = w[0], U[:, 0]
    r = o.build_gf_normal(nup, ndw, gs, e0, 1, lmats=64, lreal=16, ngfiter=200)
    wm = r["wm"]
    g = np.zeros(len(wm), dtype=complex)
    for add, isign in ((True, 1), (False, -1)):
        vv, n2, (jn, jd) = o.gf_start_vector(nup, ndw, gs, 1, 1, add)
        with o.sector(jn, jd) as sj:
            wj, Uj = np.linalg.eigh(sj.hmat())
        amp = (Uj.T @ vv) ** 2 * n2
        for k in range(len(wj)):
            g += amp[k] / (1j * wm - isign * (wj[k] - e0))
    assert np.abs(g - r["gmats"]).max() < 1e-9


def test_golden_fixture_c1():
    """tests/golden/c1_golden.npz is produced by tests/golden/make_golden.py from this oracle;
    it freezes the oracle's outputs so that later edits cannot drift silently."""
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "c1_golden.npz")
    gold = np.load(path)
    cfg, o = make_oracle("C1")
    with o.sector(4, 4) as s:
        assert np.array_equal(s.map_up(), gold["map_up"])
        rp, cols, vals = s.hup()
        assert np.array_equal(rp, gold["hup_rowptr"]) and np.array_equal(cols, gold["hup_cols"])
        assert np.array_equal(vals, gold["hup_vals"])
        assert np.array_equal(s.h0d(), gold["h0d"])
        v = configs.bench_vector(s.dim)
        v /= np.linalg.norm(v)
        assert np.abs(s.spmatvec(v) - gold["hv"]).max() < 1e-15
        e0, gs, a, b = s.lanc_eigh(v0=np.ones(s.dim) / np.sqrt(s.dim))
        assert abs(e0 - gold["e0"]) < 1e-13
        assert np.abs(a - gold["alanc"]).max() < 1e-9


def test_observables_known_answers():
    """lanc_observables / lanc_local_energy restatement (ED_OBSERVABLES.f90:95-363, 372-600) pinned on physics:
    at U = 0 the ground state is a product of the two spin species, so <n_up n_dw> = <n_up><n_dw> = 1/4 at half
    filling; with U > 0 the double occupancy drops; probabilities sum to one; the impurity density matrix has the
    densities on its diagonal; <H> recomposed from the pieces the routine returns plus the bath / hybridisation part
    equals E0 (checked through <gs|H|gs> = E0 and the diagonal identity below)."""
    for u, lo, hi in [(0.0, 0.25 - 1e-9, 0.25 + 1e-9), (2.0, 0.05, 0.2499), (8.0, 0.0, 0.06)]:
        cfg = configs.config("C1")
        cfg["uloc"] = (u,)
        o = O.Oracle(**configs.solver_kwargs(cfg))
        with o.sector(4, 4) as s:
            e0, gs, _, _ = s.lanc_eigh(v0=np.ones(s.dim) / np.sqrt(s.dim))
            assert abs(gs @ s.spmatvec(gs) - e0) < 1e-9
        ob = o.observables(4, 4, gs)
        assert abs(ob["dens"][0] - 1.0) < 1e-8 and abs(ob["dens_up"][0] - 0.5) < 1e-8
        assert lo <= ob["docc"][0] <= hi, (u, ob["docc"][0])
        assert abs(ob["prob"].sum() - 1.0) < 1e-9
        assert abs(ob["dm"][0][0] - ob["dens_up"][0]) < 1e-12
        assert abs(ob["magz"][0]) < 1e-8 and abs(ob["sz2"][0] - (ob["dens"][0] - 2 * ob["docc"][0]) / 4) < 1e-9
        # HFMODE: Epot = U<n_up n_dw> + Ehartree, Ehartree = -U/2 <n> + U/4
        assert abs(ob["ehartree"] - (-0.5 * u * ob["dens"][0] + 0.25 * u)) < 1e-8
        assert abs(ob["epot"] - (u * ob["docc"][0] + ob["ehartree"])) < 1e-9
    # two orbitals with Hund's coupling: zeta (degeneracy) only rescales, spin-exchange / pair-hopping correlators exist
    cfg, o = make_oracle("C4")
    with o.sector(5, 5) as s:
        e0, gs, _, _ = s.lanc_eigh(v0=np.ones(s.dim) / np.sqrt(s.dim))
    a, b = o.observables(5, 5, gs, zeta=1.0), o.observables(5, 5, gs, zeta=2.0)
    assert np.allclose(a["dens"], 2 * b["dens"]) and abs(a["dse"] - 2 * b["dse"]) < 1e-14
    assert a["dse"] < 0 and a["dph"] < 0 and abs(a["dens"][:2].sum() - 2.0) < 1e-7
    assert abs(a["dust"] + a["dund"] - (a["n2"][1] + a["n2"][5]) / 2) < 1e-9          # sum_{i<j} n_i n_j = Dust + Dund


def test_sparse_column_block_and_sharded_cpu_paths():
    """spmatvec_block_cols (the spot-check form that needs only the touched columns) is bit-identical to
    spmatvec_block, and the all-ranks spMatVec_MPI_main with both transposes equals the serial routine."""
    cfg, o = make_oracle("NS10V")
    with o.sector(5, 5) as full:
        v = configs.bench_vector(full.dim)
        ref = full.spmatvec(v)
        du = full.dimup
    for blk in (0, 100, 251):
        with o.sector(5, 5, blk, 252) as b:
            ci = b.block_columns()
            xc = np.stack([v[j * du:(j + 1) * du] for j in ci])
            assert np.array_equal(b.spmatvec_block_cols(ci, xc), b.spmatvec_block(v))
            assert np.abs(b.spmatvec_block(v) - ref[b.ishift:b.ishift + b.nloc]).max() < 1e-13
    secs = [o.sector(5, 5, r, 5) for r in range(5)]
    hv = np.empty_like(v)
    O.spmatvec_mpi_prebuilt(secs, v, hv, 3)
    assert np.abs(hv - ref).max() < 1e-12 * np.abs(ref).max()
    for s in secs:
        s.close()


def test_chi_chains_known_answers():
    """Susceptibility chains restated from ED_GF_CHISPIN.f90:114-488 / ED_GF_CHIDENS.f90:111-489, pinned on physics
    and on a dense spectral sum: chi(tau = 0) = norm2 = <gs|O^2|gs>, i.e. <Sz^2> = (<n> - 2<n_up n_dw>)/4 for the spin
    chain of one orbital and <n^2> = <n> + 2<n_up n_dw> for the density chain (numbers from the observables
    restatement); the mixed chain of two orbitals is norm2(i) + norm2(j) + 2<O_i O_j>; and the Lanczos chi(tau) equals
    sum_n |<n|O|gs>|^2 exp(-tau (E_n - E_0)) from a full diagonalisation of the sector."""
    cfg, o = make_oracle("NS6")
    nup = ndw = 3
    with o.sector(nup, ndw) as s:
        h = s.hmat()
        w, v = np.linalg.eigh(h)
        gs, e0 = v[:, 0], w[0]
        mu, md, du = s.map_up(), s.map_dw(), s.dimup
    ob = o.observables(nup, ndw, gs)
    tau = np.linspace(0.0, 3.0, 7)
    vm = np.pi / 50.0 * 2 * np.arange(5)
    vr = np.linspace(-2, 2, 5)
    for kind, expect in [(0, (ob["dens"][0] - 2 * ob["docc"][0]) / 4), (1, ob["dens"][0] + 2 * ob["docc"][0])]:
        vv, n2 = o.chi_start_vector(nup, ndw, gs, kind, 1, 1)
        assert abs(n2 - expect) < 1e-10 and abs(np.linalg.norm(vv) - 1) < 1e-12
        ch = o.chi_chain(nup, ndw, gs, kind, 1, 1, ngfiter=150)
        assert abs(ch["norm2"] - n2) < 1e-13
        civ, ctau, cw = O.add_to_lanczos_chi(ch["norm2"], e0, 50.0, ch["alanc"], ch["blanc"], vm, tau, vr, 0.01)
        assert abs(ctau[0] - n2) < 1e-10
        # dense spectral sum with the operator built from the bit maps
        nu = (mu[np.arange(h.shape[0]) % du] & 1).astype(float)
        nd = (md[np.arange(h.shape[0]) // du] & 1).astype(float)
        op = 0.5 * (nu - nd) if kind == 0 else nu + nd
        amp2 = (v.T @ (op * gs)) ** 2
        ref = np.array([(amp2 * np.exp(-t * (w - e0))).sum() for t in tau])
        assert np.abs(ctau - ref).max() < 1e-8
    # total == the single orbital for Norb = 1; two-orbital model: mixed chain norm
    vv1, n1 = o.chi_start_vector(nup, ndw, gs, 0, 1, 1)
    vv0, n0 = o.chi_start_vector(nup, ndw, gs, 0, 0, 0)
    assert np.array_equal(vv0, vv1) and n0 == n1
    cfg4, o4 = make_oracle("C4")
    with o4.sector(5, 5) as s:
        e4, g4, _, _ = s.lanc_eigh(v0=np.ones(s.dim) / np.sqrt(s.dim))
    _, na = o4.chi_start_vector(5, 5, g4, 1, 1, 1)
    _, nb = o4.chi_start_vector(5, 5, g4, 1, 2, 2)
    _, nm = o4.chi_start_vector(5, 5, g4, 1, 1, 2)
    _, nt = o4.chi_start_vector(5, 5, g4, 1, 0, 0)
    assert abs(nm - nt) < 1e-12 and nm > na and nm > nb            # n_1 + n_2 is the total for two orbitals


def test_orbs_operator_known_answers():
    """ed_total_ud = F restated from ed_buildh_orbs / spMatVec_orbs (ED_HAMILTONIAN_SPARSE_HxV.f90:206-370, 487-564,
    stored/Orbs/*.f90), pinned on the independently checked ed_total_ud = T path: without inter-orbital hopping and with
    Jx = Jp = 0 the total (Nup, Ndw) sector is block diagonal in the per-orbital occupations, so the union of the spectra
    of all orbital-resolved sectors with the same totals IS the spectrum of the total sector (400 eigenvalues at
    Norb = 2, Nbath = 2).  Plus: hermiticity, the sector numbering of get_Sector with QN = [Nups, Ndws], and the U = 0 /
    Ust = Jh = 0 limit where the ground-state energy is the sum of the orbitals' single-band energies."""
    cfg, o = make_oracle("ORB2")
    nt, nso = cfg["nup"], cfg["nbath"] + 1
    with o.sector(nt, nt) as s:
        w = np.linalg.eigvalsh(s.hmat())
    allw, secs = [], set()
    for nu1 in range(nso + 1):
        for nd1 in range(nso + 1):
            nu2, nd2 = nt - nu1, nt - nd1
            if not (0 <= nu2 <= nso and 0 <= nd2 <= nso):
                continue
            with o.sector_orbs([nu1, nu2], [nd1, nd2]) as so:
                hm = so.hmat()
                assert np.abs(hm - hm.T).max() < 1e-14
                allw += list(np.linalg.eigvalsh(hm))
                assert so.isector == 1 + nd2 + (nso + 1) * (nd1 + (nso + 1) * (nu2 + (nso + 1) * nu1))
                secs.add(so.isector)
                assert so.dim == int(np.prod(so.dims)) and so.dims[0] == O.lib().orc_binomial(nso, nu1)
    assert len(allw) == len(w) and np.abs(np.sort(allw) - w).max() < 1e-12
    # decoupled orbitals: E0(orbs sector) = E0(orbital 1 alone) + E0(orbital 2 alone)
    kw = configs.solver_kwargs(cfg)
    kw.update(ust=0.0, jh=0.0)
    o2 = O.Oracle(**kw)
    with o2.sector_orbs([1, 2], [2, 1]) as so:
        e_pair = np.linalg.eigvalsh(so.hmat())[0]
    e_sum = 0.0
    for io, (nu, nd) in enumerate([(1, 2), (2, 1)]):
        k1 = dict(norb=1, nbath=cfg["nbath"], nspin=1, uloc=(cfg["uloc"][io],), ust=0.0, jh=0.0, jx=0.0, jp=0.0, xmu=cfg["xmu"],
                  hfmode=True, bath_e=np.asarray(cfg["bath_e"])[:, io:io + 1, :], bath_v=np.asarray(cfg["bath_v"])[:, io:io + 1, :])
        with O.Oracle(**k1).sector(nu, nd) as s1:
            e_sum += np.linalg.eigvalsh(s1.hmat())[0]
    assert abs(e_pair - e_sum) < 1e-12


def test_phonon_operator_known_answers():
    """spMatVec_main with DimPh > 1 (ED_HAMILTONIAN_SPARSE_HxV.f90:391-485, stored/H_ph.f90, H_e_ph.f90) against the
    Kronecker construction H = 1_ph x H_el + diag(w0 n) x 1_el + (b + b^+) x diag(sum_orb g (n_up + n_dw - 1)) built with
    numpy from the electron sector's dense matrix and the basis maps; g = 0 decouples (spectrum = E_el + w0 n)."""
    cfg, o = make_oracle("NS6")
    nph, g, w0 = 3, (0.7,), 0.4
    with o.sector(3, 3) as s:
        hel = s.hmat()
        mu, md, du, dim = s.map_up(), s.map_dw(), s.dimup, s.dim
        nimp = (mu[np.arange(dim) % du] & 1) + (md[np.arange(dim) // du] & 1)
        e = g[0] * (nimp - 1.0)
        b = np.diag(np.sqrt(np.arange(1, nph + 1)), 1)                      # destruction operator
        hfull = np.kron(np.eye(nph + 1), hel) + np.kron(np.diag(w0 * np.arange(nph + 1)), np.eye(dim)) + np.kron(b + b.T, np.diag(e))
        rng = np.random.default_rng(3)
        v = rng.standard_normal(dim * (nph + 1))
        assert np.abs(s.spmatvec_ph(v, nph, g, w0) - hfull @ v).max() < 1e-12
        w_el = np.linalg.eigvalsh(hel)
        v0 = np.ones(dim * (nph + 1)) / np.sqrt(dim * (nph + 1))
        e0, vec, _, _ = s.lanc_eigh_ph(nph, g, w0, v0)
        assert abs(e0 - np.linalg.eigvalsh(hfull)[0]) < 1e-10
        e00, _, _, _ = s.lanc_eigh_ph(nph, (0.0,), w0, v0)
        assert abs(e00 - w_el[0]) < 1e-10                                   # g = 0: the phonon vacuum on top of the electron ground state
        assert e0 < e00                                                     # the coupling lowers the energy (second order)


def test_hybrid_and_replica_baths_known_answers():
    """bath_type hybrid and replica (getBathStride ED_SETUP.f90:358-375, the replica inter-orbital bath hopping of
    stored/H_up.f90:26-50, bath_diag / diag_hybr of ed_buildh_main :46-76) pinned on (1) a replica bath with DIAGONAL
    Hbath matrices is the normal bath with e(a,k) = Hbath(a,a,k), V(a,k) = v(k) up to a relabelling of the sites: identical
    spectra, interaction, spin-exchange and pair-hopping included; (2) at U = 0 the ground-state energy of any sector is
    the sum of the lowest one-body levels of the Ns x Ns hopping matrix, for a replica bath with off-diagonal Hbath and
    for a hybrid bath."""
    norb, nbath = 2, 2
    rng = np.random.default_rng(1)
    hb = np.zeros((1, 1, norb, norb, nbath))
    for kp in range(nbath):
        for io in range(norb):
            hb[0, 0, io, io, kp] = rng.standard_normal()
    v = np.array([[0.4, 0.7]])
    kw = dict(norb=norb, nbath=nbath, nspin=1, uloc=(2.0, 1.5), ust=0.8, jh=0.3, jx=0.3, jp=0.3, xmu=0.2, hfmode=True)
    orep = O.Oracle(bath_type=2, bath_v=v, bath_h=hb, **kw)
    e = np.zeros((1, norb, nbath))
    vv = np.zeros((1, norb, nbath))
    for io in range(norb):
        for kp in range(nbath):
            e[0, io, kp], vv[0, io, kp] = hb[0, 0, io, io, kp], v[0, kp]
    onor = O.Oracle(bath_e=e, bath_v=vv, **kw)
    for sec in [(3, 3), (2, 4)]:
        with orep.sector(*sec) as a, onor.sector(*sec) as b:
            assert np.abs(np.linalg.eigvalsh(a.hmat()) - np.linalg.eigvalsh(b.hmat())).max() < 1e-12
    hb2 = hb.copy()
    hb2[0, 0, 0, 1, :] = hb2[0, 0, 1, 0, :] = [0.3, -0.2]
    kw0 = dict(norb=norb, nbath=nbath, nspin=1, uloc=(0.0, 0.0), ust=0.0, jh=0.0, jx=0.0, jp=0.0, xmu=0.0, hfmode=False)
    o0 = O.Oracle(bath_type=2, bath_v=v, bath_h=hb2, **kw0)
    ns = norb * (nbath + 1)
    h1 = np.zeros((ns, ns))
    for kp in range(nbath):
        for io in range(norb):
            sb = io + (kp + 1) * norb
            h1[io, sb] = h1[sb, io] = v[0, kp]
            for jo in range(norb):
                h1[sb, jo + (kp + 1) * norb] = hb2[0, 0, io, jo, kp]
    w1 = np.linalg.eigvalsh(h1)
    for sec in [(3, 3), (2, 1)]:
        with o0.sector(*sec) as a:
            assert abs(np.linalg.eigvalsh(a.hmat())[0] - (w1[:sec[0]].sum() + w1[:sec[1]].sum())) < 1e-12
    vh = np.array([[[0.5, 0.3, 0.2], [0.1, 0.4, 0.6]]])
    eh = np.array([[[-1.0, 0.2, 0.9]]])
    oh = O.Oracle(norb=2, nbath=3, nspin=1, uloc=(0.0, 0.0), ust=0.0, jh=0.0, jx=0.0, jp=0.0, xmu=0.0, hfmode=False,
                  bath_type=1, bath_e=eh, bath_v=vh)
    h1 = np.zeros((5, 5))
    for kp in range(3):
        h1[2 + kp, 2 + kp] = eh[0, 0, kp]
        for io in range(2):
            h1[io, 2 + kp] = h1[2 + kp, io] = vh[0, io, kp]
    w1 = np.linalg.eigvalsh(h1)
    for sec in [(2, 3), (1, 1)]:
        with oh.sector(*sec) as a:
            assert abs(np.linalg.eigvalsh(a.hmat())[0] - (w1[:sec[0]].sum() + w1[:sec[1]].sum())) < 1e-12


def test_noninteracting_gf_is_the_hybridisation_formula():
    """U = 0: the impurity Green's function of the whole chain (ground state by Lanczos, c / c^+ start vectors, GF
    Lanczos chains, add_to_lanczos_gf_normal) must equal 1 / (iw + mu - e_imp - sum_k V_k^2 / (iw - e_k)) -- the
    closed form behind build_sigma_normal (ED_GF_NORMAL.f90:935-1002, ED_BATH_FUNCTIONS.f90:163-195) -- so Sigma
    vanishes identically.  Ties the sign / site conventions of the operators, the chains' weights and poles, and the
    bath formula together without any reference output."""
    cfg = configs.config("NS6V")                                   # level-dependent V_k, asymmetric levels, xmu != 0
    cfg["uloc"] = (0.0,)
    cfg["hfmode"] = False
    o = O.Oracle(**configs.solver_kwargs(cfg))
    # ground state over all sectors at U = 0: fill the lowest one-body levels of each spin
    ns = cfg["nbath"] + 1
    best = None
    for nup in range(ns + 1):
        for ndw in range(ns + 1):
            with o.sector(nup, ndw) as s:
                w, v = np.linalg.eigh(s.hmat())
            if best is None or w[0] < best[0] - 1e-12:
                best = (w[0], v[:, 0], nup, ndw)
    e0, gs, nup, ndw = best
    r = o.build_gf_normal(nup, ndw, gs, e0, 1, lmats=64, lreal=8, ngfiter=200)
    wm = r["wm"]
    e = np.asarray(cfg["bath_e"]).reshape(-1)
    v = np.asarray(cfg["bath_v"]).reshape(-1)
    z = 1j * wm
    delta = (v[None, :] ** 2 / (z[:, None] - e[None, :])).sum(axis=1)
    g_exact = 1.0 / (z + cfg["xmu"] - delta)
    assert np.abs(r["gmats"] - g_exact).max() < 1e-9
    sig, _ = o.sigma_normal(1, 1, z, r["gmats"])
    assert np.abs(sig).max() < 1e-8


def test_golden_fixture_round2():
    """tests/golden/round2_golden.npz (tests/golden/make_golden.py::round2) freezes the oracle's outputs for the round-2
    additions -- susceptibility chains, orbital-resolved operator, phonon terms, replica and hybrid baths -- so that later
    edits of the restatement cannot drift silently: integer structure bit-exact, values to rounding."""
    import os
    import runpy
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "round2_golden.npz"))
    mod = runpy.run_path(os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        mod["round2"].__globals__["HERE"] = td                      # regenerate next to nothing, compare key by key
        mod["round2"]()
        new = np.load(os.path.join(td, "round2_golden.npz"))
        assert set(new.files) == set(gold.files)
        for k in gold.files:
            a, b = gold[k], new[k]
            if np.issubdtype(a.dtype, np.integer):
                assert np.array_equal(a, b), k
            elif k.endswith("_vals") or k.endswith("_h0d"):
                assert np.array_equal(a, b), k                      # matrix elements and diagonals are deterministic sums
            elif k.endswith("anc"):
                assert np.abs(a[:10] - b[:10]).max() < 1e-10, k     # Lanczos coefficients: the stable prefix
            else:
                assert np.abs(a - b).max() <= 1e-12 * max(1.0, np.abs(a).max()), k
