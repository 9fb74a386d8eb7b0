"""ctypes front-end of the CPU ORACLE (oracle/libedoracle.so).

TEST INFRASTRUCTURE ONLY -- parity unpinned (see oracle/ed_oracle.h).  Import this module only
from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)


def build(force=False):
    """Compile oracle/libedoracle.so with the committed Makefile (gcc only)."""
    so = os.path.join(_HERE, "libedoracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("ed_oracle.c", "ed_oracle.h", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libedoracle.so"])
    return so


class _Csr(C.Structure):
    _fields_ = [("nrow", C.c_int64), ("ncol", C.c_int64), ("rowptr", c_i64p),
                ("cols", c_i64p), ("vals", c_dp)]


class Observables(C.Structure):
    """orc_observables / edgpu_observables (same layout): ED_OBSERVABLES.f90 outputs of one state."""
    _fields_ = [("dens", C.c_double * 5), ("dens_up", C.c_double * 5), ("dens_dw", C.c_double * 5), ("docc", C.c_double * 5),
                ("magz", C.c_double * 5), ("sz2", C.c_double * 25), ("n2", C.c_double * 25), ("s2tot", C.c_double),
                ("prob", C.c_double * 243), ("dm", (C.c_double * 25) * 2),
                ("eknot", C.c_double), ("epot", C.c_double), ("ehartree", C.c_double), ("dust", C.c_double),
                ("dund", C.c_double), ("dse", C.c_double), ("dph", C.c_double)]

    def as_dict(self):
        out = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            out[name] = float(v) if isinstance(v, float) else np.array(v, dtype=np.float64)
        return out


class _Sector(C.Structure):
    _fields_ = [("ctx", C.c_void_p), ("nup", C.c_int), ("ndw", C.c_int),
                ("dimup", C.c_int64), ("dimdw", C.c_int64), ("dim", C.c_int64),
                ("map_up", c_i32p), ("map_dw", c_i32p),
                ("rank", C.c_int), ("nranks", C.c_int),
                ("qdw", C.c_int64), ("rdw", C.c_int64), ("q", C.c_int64), ("r", C.c_int64),
                ("istart", C.c_int64), ("iend", C.c_int64), ("ishift", C.c_int64),
                ("sparse_h", C.c_int), ("h0d", c_dp),
                ("hup", _Csr), ("hdw", _Csr), ("hnd", _Csr)]


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libedoracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_ctx_create.restype = C.c_void_p
        L.orc_ctx_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, c_dp, C.c_double, C.c_double,
                                     C.c_double, C.c_double, C.c_double, c_dp, c_dp, c_dp]
        L.orc_ctx_create_bt.restype = C.c_void_p
        L.orc_ctx_create_bt.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, c_dp, C.c_double, C.c_double,
                                        C.c_double, C.c_double, C.c_double, c_dp, C.c_int, c_dp, c_dp, c_dp]
        L.orc_ctx_destroy.argtypes = [C.c_void_p]
        L.orc_init_dmft_bath.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, c_dp, c_dp]
        L.orc_binomial.restype = C.c_int
        L.orc_build_sector_map.restype = C.c_int64
        L.orc_build_sector_map.argtypes = [C.c_int, C.c_int, c_i32p]
        L.orc_c.argtypes = [C.c_int, C.c_int32, c_i32p, c_dp]
        L.orc_cdg.argtypes = [C.c_int, C.c_int32, c_i32p, c_dp]
        L.orc_binary_search.restype = C.c_int64
        L.orc_binary_search.argtypes = [c_i32p, C.c_int64, C.c_int32]
        L.orc_build_hv_sector.restype = C.POINTER(_Sector)
        L.orc_build_hv_sector.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_delete_hv_sector.argtypes = [C.POINTER(_Sector)]
        L.orc_vecdim_hv_sector.restype = C.c_int64
        L.orc_vecdim_hv_sector.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_build_hmat.argtypes = [C.POINTER(_Sector), c_dp]
        L.orc_spmatvec_main.argtypes = [C.POINTER(_Sector), C.c_int64, c_dp, c_dp]
        L.orc_directmatvec_main.argtypes = [C.POINTER(_Sector), C.c_int64, c_dp, c_dp]
        L.orc_spmatvec_block.argtypes = [C.POINTER(_Sector), c_dp, c_dp]
        L.orc_spmatvec_block_cols.argtypes = [C.POINTER(_Sector), C.POINTER(C.c_int64), C.c_int64, c_dp, c_dp]
        L.orc_spmatvec_blocks_mt.restype = C.c_int64
        L.orc_spmatvec_blocks_mt.argtypes = [C.POINTER(C.POINTER(_Sector)), C.c_int, C.c_int, c_dp, C.POINTER(c_dp)]
        L.orc_spmatvec_mpi_main_all.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_dp, c_dp]
        L.orc_directmatvec_mpi_main_all.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_dp, c_dp]
        L.orc_spmatvec_mpi_main_prebuilt.argtypes = [C.POINTER(C.POINTER(_Sector)), C.c_int, C.c_int, c_dp, c_dp]
        L.orc_vector_transpose_all.argtypes = [C.c_int, C.c_int64, C.c_int64, C.POINTER(c_dp), C.POINTER(c_dp)]
        L.orc_tql2.argtypes = [C.c_int, c_dp, c_dp, c_dp]
        L.orc_lanc_eigh_sector.argtypes = [C.POINTER(_Sector), C.c_int, c_dp, c_dp, C.c_int, C.c_double,
                                           C.c_int, c_ip, c_dp, c_dp]
        L.orc_lanc_tridiag_sector.argtypes = [C.POINTER(_Sector), C.c_int, c_dp, c_dp, c_dp, C.c_int, C.c_double]
        L.orc_gf_start_vector.restype = C.c_int64
        L.orc_gf_start_vector.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp, C.c_int, C.c_int, C.c_int,
                                          c_dp, c_dp, c_ip, c_ip]
        L.orc_add_to_lanczos_gf.argtypes = [C.c_double, C.c_double, C.c_double, c_dp, c_dp, C.c_int, C.c_int,
                                            c_dp, C.c_int, c_dp, c_dp, C.c_int, C.c_double, c_dp]
        L.orc_lanc_build_gf_normal_main.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp, C.c_double, C.c_double,
                                                    C.c_int, C.c_int, C.c_int, C.c_int, c_dp, C.c_int, c_dp,
                                                    c_dp, C.c_int, C.c_double, c_dp, c_dp, c_ip]
        L.orc_observables_normal.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp, C.c_double, C.POINTER(Observables)]
        L.orc_sigma_normal.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp, C.c_int, c_dp, c_dp, c_dp]
        L.orc_allocate_grids.argtypes = [C.c_double, C.c_int, C.c_double, C.c_double, C.c_int, c_dp, c_dp]
        L.orc_spmatvec_main_ph.argtypes = [C.POINTER(_Sector), C.c_int, c_dp, C.c_double, C.c_int64, c_dp, c_dp]
        L.orc_lanc_eigh_sector_ph.argtypes = [C.POINTER(_Sector), C.c_int, c_dp, C.c_double, c_dp, c_dp, C.c_int, C.c_double, C.c_int,
                                              c_ip, c_dp, c_dp]
        L.orc_get_sector_orbs.argtypes = [C.c_void_p, c_ip, c_ip]
        L.orc_build_hv_sector_orbs.restype = C.c_void_p
        L.orc_build_hv_sector_orbs.argtypes = [C.c_void_p, c_ip, c_ip]
        L.orc_delete_hv_sector_orbs.argtypes = [C.c_void_p]
        L.orc_spmatvec_orbs.argtypes = [C.c_void_p, C.c_int64, c_dp, c_dp]
        L.orc_lanc_eigh_sector_orbs.argtypes = [C.c_void_p, c_dp, c_dp, C.c_int, C.c_double, C.c_int, c_ip, c_dp, c_dp]
        L.orc_lanc_tridiag_sector_orbs.argtypes = [C.c_void_p, c_dp, c_dp, c_dp, C.c_int, C.c_double]
        L.orc_sector_orbs_info.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_sector_orbs_get.argtypes = [C.c_void_p, C.c_int, c_i32p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), c_dp, c_dp]
        L.orc_chi_start_vector.restype = C.c_int64
        L.orc_chi_start_vector.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp, C.c_int, C.c_int, C.c_int, c_dp, c_dp]
        L.orc_chi_chain.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_dp, c_dp, c_dp]
        L.orc_add_to_lanczos_chi.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, c_dp, c_dp, C.c_int,
                                             c_dp, C.c_int, c_dp, c_dp, C.c_int, c_dp, c_dp, C.c_int, C.c_double, c_dp]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def init_dmft_bath(norb, nbath, nspin=1, hwband=2.0):
    """Returns (e, v) shaped (nspin, norb, nbath) [ED_BATH/dmft_aux.f90:102-133]."""
    e = np.zeros(nspin * norb * nbath)
    v = np.zeros(nspin * norb * nbath)
    lib().orc_init_dmft_bath(norb, nbath, nspin, hwband, _dp(e), _dp(v))
    shp = (nspin, norb, nbath)
    return e.reshape(shp, order="F"), v.reshape(shp, order="F")


def build_sector_map(ns, n):
    dim = lib().orc_build_sector_map(ns, n, None)
    m = np.zeros(dim, dtype=np.int32)
    lib().orc_build_sector_map(ns, n, m.ctypes.data_as(c_i32p))
    return m


def _csr_np(cs):
    n = cs.nrow
    rowptr = np.ctypeslib.as_array(cs.rowptr, shape=(n + 1,)).copy() if n else np.zeros(1, np.int64)
    nnz = int(rowptr[-1])
    cols = np.ctypeslib.as_array(cs.cols, shape=(nnz,)).copy() if nnz else np.zeros(0, np.int64)
    vals = np.ctypeslib.as_array(cs.vals, shape=(nnz,)).copy() if nnz else np.zeros(0)
    return rowptr, cols, vals


class Sector:
    """build_Hv_sector ... delete_Hv_sector lifetime (ED_HAMILTONIAN.f90:43-222)."""

    def __init__(self, ora, nup, ndw, rank=0, nranks=1, sparse_h=True):
        self.ora = ora
        self.p = lib().orc_build_hv_sector(ora.h, nup, ndw, rank, nranks, int(bool(sparse_h)))
        s = self.p.contents
        self.nup, self.ndw = nup, ndw
        self.dimup, self.dimdw, self.dim = s.dimup, s.dimdw, s.dim
        self.qdw, self.ishift, self.istart, self.iend = s.qdw, s.ishift, s.istart, s.iend
        self.nloc = s.dimup * s.qdw
        self.sparse_h = bool(sparse_h)

    def close(self):
        if self.p is not None:
            lib().orc_delete_hv_sector(self.p)
            self.p = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # introspection
    def map_up(self):
        return np.ctypeslib.as_array(self.p.contents.map_up, shape=(self.dimup,)).copy()

    def map_dw(self):
        return np.ctypeslib.as_array(self.p.contents.map_dw, shape=(self.dimdw,)).copy()

    def h0d(self):
        return np.ctypeslib.as_array(self.p.contents.h0d, shape=(self.nloc,)).copy()

    def hup(self):
        return _csr_np(self.p.contents.hup)

    def hdw(self):
        return _csr_np(self.p.contents.hdw)

    def hnd(self):
        return _csr_np(self.p.contents.hnd)

    def hmat(self):
        h = np.zeros(self.dim * self.dim)
        lib().orc_build_hmat(self.p, _dp(h))
        return h.reshape((self.dim, self.dim), order="F")

    # operators
    def spmatvec(self, v):
        v = _f64(v)
        hv = np.empty_like(v)
        lib().orc_spmatvec_main(self.p, v.size, _dp(v), _dp(hv))
        return hv

    def spmatvec_block(self, v_full):
        """spMatVec_main loops for this sector view's column block only (full vector in)."""
        v = _f64(v_full)
        hv = np.empty(self.nloc)
        lib().orc_spmatvec_block(self.p, _dp(v), _dp(hv))
        return hv

    def block_columns(self):
        """Ascending global i_dw indices of the columns of v that spmatvec_block reads for this block."""
        rp, cols, _ = self.hdw()
        c0, c1 = self.istart // self.dimup, self.iend // self.dimup
        need = set(range(c0, c1))
        for j in range(c0, c1):
            need.update(int(x) for x in cols[rp[j]:rp[j + 1]])
        return np.array(sorted(need), dtype=np.int64)

    def spmatvec_block_cols(self, colidx, xcols):
        """spmatvec_block for a caller that holds only the needed columns (xcols[k] = column colidx[k])."""
        colidx = np.ascontiguousarray(colidx, dtype=np.int64)
        x = _f64(np.ascontiguousarray(xcols).reshape(-1))
        assert x.size == colidx.size * self.dimup
        hv = np.empty(self.nloc)
        rc = lib().orc_spmatvec_block_cols(self.p, colidx.ctypes.data_as(C.POINTER(C.c_int64)), colidx.size, _dp(x), _dp(hv))
        assert rc == 0
        return hv

    def spmatvec_ph(self, v, nph, g_ph, w0_ph):
        """spMatVec_main with DimPh = nph + 1 phonon slabs (v of length dim*(nph+1), phonon index slowest)."""
        v = _f64(v)
        g = _f64(list(g_ph) + [0.0] * (5 - len(g_ph)))
        hv = np.zeros(self.dim * (nph + 1))
        lib().orc_spmatvec_main_ph(self.p, nph, _dp(g), C.c_double(w0_ph), v.size, _dp(v), _dp(hv))
        return hv

    def lanc_eigh_ph(self, nph, g_ph, w0_ph, v0, nitermax=512, threshold=1e-18, ncheck=10):
        vect = _f64(v0).copy()
        g = _f64(list(g_ph) + [0.0] * (5 - len(g_ph)))
        nit = min(vect.size, nitermax)
        egs = C.c_double(0.0)
        nl = C.c_int(0)
        a = np.zeros(nit + 1)
        b = np.zeros(nit + 1)
        lib().orc_lanc_eigh_sector_ph(self.p, nph, _dp(g), C.c_double(w0_ph), C.byref(egs), _dp(vect), nit, threshold, ncheck,
                                      C.byref(nl), _dp(a), _dp(b))
        return egs.value, vect, a[:nl.value].copy(), b[:nl.value].copy()

    def directmatvec(self, v):
        v = _f64(v)
        hv = np.empty_like(v)
        lib().orc_directmatvec_main(self.p, v.size, _dp(v), _dp(hv))
        return hv

    def lanc_eigh(self, v0=None, nitermax=512, threshold=1e-18, ncheck=10, mode=0):
        """sp_lanc_eigh as called at ED_DIAG.f90:177-185.  Returns (egs, vect, alanc, blanc)."""
        nit = int(min(self.dim, nitermax))
        vect = np.zeros(self.dim) if v0 is None else _f64(v0).copy()
        egs = C.c_double(0.0)
        nl = C.c_int(0)
        a = np.zeros(nit + 2)
        b = np.zeros(nit + 2)
        lib().orc_lanc_eigh_sector(self.p, mode, C.byref(egs), _dp(vect), nit, threshold, ncheck,
                                   C.byref(nl), _dp(a), _dp(b))
        return egs.value, vect, a[:nl.value].copy(), b[:nl.value].copy()

    def lanc_tridiag(self, vin, nlanc, threshold=1e-12, mode=0):
        """sp_lanc_tridiag as called at ED_GF_NORMAL.f90:232-237.  Returns (alanc, blanc)."""
        v = _f64(vin).copy()
        a = np.zeros(nlanc)
        b = np.zeros(nlanc)
        lib().orc_lanc_tridiag_sector(self.p, mode, _dp(v), _dp(a), _dp(b), nlanc, threshold)
        return a, b


class OrbsSector:
    """build_Hv_sector for ed_total_ud = F (one (Nup, Ndw) per orbital): ed_buildh_orbs + spMatVec_orbs."""

    def __init__(self, ora, nups, ndws):
        self.ora = ora
        self.norb = len(nups)
        nu = (C.c_int * self.norb)(*nups)
        nd = (C.c_int * self.norb)(*ndws)
        self.isector = lib().orc_get_sector_orbs(ora.h, nu, nd)
        self.s = lib().orc_build_hv_sector_orbs(ora.h, nu, nd)
        dim = C.c_int64(0)
        dims = (C.c_int64 * (2 * self.norb))()
        lib().orc_sector_orbs_info(self.s, C.byref(dim), dims)
        self.dim, self.dims = dim.value, list(dims)

    def close(self):
        if self.s:
            lib().orc_delete_hv_sector_orbs(self.s)
            self.s = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def factor(self, f):
        """(map, rowptr, cols, vals) of factor f: up word of orbital f+1 for f < Norb, dw word of orbital f+1-Norb else."""
        n = self.dims[f]
        m = np.zeros(n, np.int32)
        rp = np.zeros(n + 1, np.int64)
        lib().orc_sector_orbs_get(self.s, f, m.ctypes.data_as(c_i32p), rp.ctypes.data_as(C.POINTER(C.c_int64)), None, None, None)
        cols = np.zeros(max(int(rp[-1]), 1), np.int64)
        vals = np.zeros(max(int(rp[-1]), 1))
        lib().orc_sector_orbs_get(self.s, f, None, None, cols.ctypes.data_as(C.POINTER(C.c_int64)), _dp(vals), None)
        return m, rp, cols[:rp[-1]], vals[:rp[-1]]

    def h0d(self):
        d = np.zeros(self.dim)
        lib().orc_sector_orbs_get(self.s, 0, None, None, None, None, _dp(d))
        return d

    def spmatvec(self, v):
        v = _f64(v)
        hv = np.zeros(self.dim)
        lib().orc_spmatvec_orbs(self.s, self.dim, _dp(v), _dp(hv))
        return hv

    def hmat(self):
        h = np.zeros((self.dim, self.dim))
        e = np.zeros(self.dim)
        for j in range(self.dim):
            e[:] = 0.0
            e[j] = 1.0
            h[:, j] = self.spmatvec(e)
        return h

    def lanc_eigh(self, v0=None, nitermax=512, threshold=1e-18, ncheck=10):
        vect = np.zeros(self.dim) if v0 is None else _f64(v0).copy()
        nit = min(self.dim, nitermax)
        egs = C.c_double(0.0)
        nl = C.c_int(0)
        a = np.zeros(nit + 1)
        b = np.zeros(nit + 1)
        lib().orc_lanc_eigh_sector_orbs(self.s, C.byref(egs), _dp(vect), nit, threshold, ncheck, C.byref(nl), _dp(a), _dp(b))
        return egs.value, vect, a[:nl.value].copy(), b[:nl.value].copy()

    def lanc_tridiag(self, vin, nlanc, threshold=1e-12):
        v = _f64(vin).copy()
        a = np.zeros(nlanc)
        b = np.zeros(nlanc)
        lib().orc_lanc_tridiag_sector_orbs(self.s, _dp(v), _dp(a), _dp(b), nlanc, threshold)
        return a, b


class Oracle:
    """Holds the module-global inputs of the reference (orc_ctx)."""

    def __init__(self, norb, nbath, nspin=1, uloc=(2.0,), ust=0.0, jh=0.0, jx=0.0, jp=0.0, xmu=0.0,
                 hfmode=True, imphloc=None, bath_e=None, bath_v=None, hwband=2.0, bath_type=0, bath_h=None):
        """bath_type 0 normal: bath_e, bath_v (nspin,norb,nbath); 1 hybrid: bath_e (nspin,1,nbath), bath_v (nspin,norb,nbath);
        2 replica: bath_v (nspin,nbath), bath_h (nspin,nspin,norb,norb,nbath)."""
        self.norb, self.nbath, self.nspin = norb, nbath, nspin
        self.bath_type = int(bath_type)
        self.ns = {0: (nbath + 1) * norb, 1: norb + nbath, 2: (nbath + 1) * norb}[self.bath_type]
        ul = np.zeros(5)
        ul[:len(uloc)] = uloc
        if self.bath_type != 0:
            if imphloc is None:
                imphloc = np.zeros((nspin, nspin, norb, norb))
            self.imphloc = np.asfortranarray(imphloc, dtype=np.float64)
            fh = np.ravel(self.imphloc, order="F").copy()
            fv = np.ravel(np.asarray(bath_v, dtype=np.float64), order="F").copy()
            fe = np.ravel(np.asarray(bath_e, dtype=np.float64), order="F").copy() if bath_e is not None else np.zeros(1)
            fb = np.ravel(np.asarray(bath_h, dtype=np.float64), order="F").copy() if bath_h is not None else np.zeros(1)
            self.uloc, self.ust, self.jh, self.jx, self.jp, self.xmu, self.hfmode = ul, ust, jh, jx, jp, xmu, hfmode
            self.h = lib().orc_ctx_create_bt(norb, nbath, nspin, int(hfmode), _dp(ul), ust, jh, jx, jp, xmu, _dp(fh),
                                             self.bath_type, _dp(fe), _dp(fv), _dp(fb))
            return
        if bath_e is None or bath_v is None:
            bath_e, bath_v = init_dmft_bath(norb, nbath, nspin, hwband)
        self.bath_e = np.asfortranarray(bath_e, dtype=np.float64).reshape((nspin, norb, nbath), order="F")
        self.bath_v = np.asfortranarray(bath_v, dtype=np.float64).reshape((nspin, norb, nbath), order="F")
        if imphloc is None:
            imphloc = np.zeros((nspin, nspin, norb, norb))
        self.imphloc = np.asfortranarray(imphloc, dtype=np.float64)
        self.uloc, self.ust, self.jh, self.jx, self.jp, self.xmu, self.hfmode = ul, ust, jh, jx, jp, xmu, hfmode
        fe = np.ravel(self.bath_e, order="F").copy()
        fv = np.ravel(self.bath_v, order="F").copy()
        fh = np.ravel(self.imphloc, order="F").copy()
        self.h = lib().orc_ctx_create(norb, nbath, nspin, int(hfmode), _dp(ul), ust, jh, jx, jp, xmu,
                                      _dp(fh), _dp(fe), _dp(fv))

    def __del__(self):
        try:
            if self.h:
                lib().orc_ctx_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def sector(self, nup, ndw, rank=0, nranks=1, sparse_h=True):
        return Sector(self, nup, ndw, rank, nranks, sparse_h)

    def sector_orbs(self, nups, ndws):
        return OrbsSector(self, nups, ndws)

    def vecdim(self, nup, ndw, rank, nranks):
        return lib().orc_vecdim_hv_sector(self.h, nup, ndw, rank, nranks)

    def spmatvec_mpi(self, nup, ndw, nranks, v, nthreads=1):
        v = _f64(v)
        hv = np.empty_like(v)
        lib().orc_spmatvec_mpi_main_all(self.h, nup, ndw, nranks, nthreads, _dp(v), _dp(hv))
        return hv

    def directmatvec_mpi(self, nup, ndw, nranks, v, nthreads=1):
        v = _f64(v)
        hv = np.empty_like(v)
        lib().orc_directmatvec_mpi_main_all(self.h, nup, ndw, nranks, nthreads, _dp(v), _dp(hv))
        return hv

    def gf_start_vector(self, nup, ndw, gs, iorb, ispin, add):
        gs = _f64(gs)
        jn, jd = C.c_int(0), C.c_int(0)
        n2 = C.c_double(0.0)
        jdim = lib().orc_gf_start_vector(self.h, nup, ndw, _dp(gs), iorb, ispin, int(add), None,
                                         C.byref(n2), C.byref(jn), C.byref(jd))
        if jdim == 0:
            return None, 0.0, (jn.value, jd.value)
        vv = np.zeros(jdim)
        lib().orc_gf_start_vector(self.h, nup, ndw, _dp(gs), iorb, ispin, int(add), _dp(vv),
                                  C.byref(n2), C.byref(jn), C.byref(jd))
        return vv, n2.value, (jn.value, jd.value)

    def build_gf_normal(self, nup, ndw, gs, e0, iorb, ispin=1, ngfiter=200, mode=0, zeta=1.0,
                        beta=1000.0, lmats=5000, wini=-5.0, wfin=5.0, lreal=5000, eps=0.01):
        """lanc_build_gf_normal_main for one (iorb, ispin).  Returns dict with G, chains, grids."""
        wm, wr = allocate_grids(beta, lmats, wini, wfin, lreal)
        gm = np.zeros(lmats, dtype=np.complex128)
        gr = np.zeros(lreal, dtype=np.complex128)
        chain = np.zeros((2, 1 + 2 * ngfiter))
        nl = (C.c_int * 2)(0, 0)
        gs = _f64(gs)
        lib().orc_lanc_build_gf_normal_main(self.h, nup, ndw, _dp(gs), e0, zeta, iorb, ispin, ngfiter, mode,
                                            _dp(wm), lmats, gm.ctypes.data_as(c_dp),
                                            _dp(wr), lreal, eps, gr.ctypes.data_as(c_dp),
                                            _dp(chain), nl)
        out = {"wm": wm, "wr": wr, "gmats": gm, "greal": gr, "chains": []}
        for p in range(2):
            n = nl[p]
            out["chains"].append({"norm2": chain[p, 0], "alanc": chain[p, 1:1 + n].copy(),
                                  "blanc": chain[p, 1 + ngfiter:1 + ngfiter + n].copy(), "nlanc": n})
        return out

    def chi_start_vector(self, nup, ndw, gs, kind, iorb, jorb):
        """O|gs> normalised, and its norm2 (ED_GF_CHISPIN.f90 / ED_GF_CHIDENS.f90 start vectors)."""
        gs = _f64(gs)
        vv = np.zeros_like(gs)
        n2 = C.c_double(0.0)
        lib().orc_chi_start_vector(self.h, nup, ndw, _dp(gs), kind, iorb, jorb, _dp(vv), C.byref(n2))
        return vv, n2.value

    def chi_chain(self, nup, ndw, gs, kind, iorb, jorb, ngfiter=200, mode=0):
        """One susceptibility chain: dict(norm2, alanc, blanc, nlanc)."""
        gs = _f64(gs)
        a = np.zeros(ngfiter)
        b = np.zeros(ngfiter)
        n2 = C.c_double(0.0)
        n = lib().orc_chi_chain(self.h, nup, ndw, _dp(gs), kind, iorb, jorb, ngfiter, mode, C.byref(n2), _dp(a), _dp(b))
        return {"norm2": n2.value, "alanc": a[:n].copy(), "blanc": b[:n].copy(), "nlanc": n}

    def observables(self, nup, ndw, gs, zeta=1.0):
        """lanc_observables + lanc_local_energy of one state at T = 0 (ED_OBSERVABLES.f90:95-363, 372-600)."""
        gs = _f64(gs)
        o = Observables()
        lib().orc_observables_normal(self.h, nup, ndw, _dp(gs), zeta, C.byref(o))
        return o.as_dict()

    def sigma_normal(self, iorb, ispin, z, g):
        z = np.ascontiguousarray(z, dtype=np.complex128)
        g = np.ascontiguousarray(g, dtype=np.complex128)
        s = np.zeros_like(z)
        i0 = np.zeros_like(z)
        lib().orc_sigma_normal(self.h, iorb, ispin, z.ctypes.data_as(c_dp), z.size, g.ctypes.data_as(c_dp),
                               s.ctypes.data_as(c_dp), i0.ctypes.data_as(c_dp))
        return s, i0


def allocate_grids(beta, lmats, wini, wfin, lreal):
    wm = np.zeros(lmats)
    wr = np.zeros(lreal)
    lib().orc_allocate_grids(beta, lmats, wini, wfin, lreal, _dp(wm), _dp(wr))
    return wm, wr


def add_to_lanczos_gf(norm2, ei, alanc, blanc, isign, wm, wr, eps, zeta=1.0):
    a, b = _f64(alanc), _f64(blanc)
    gm = np.zeros(len(wm), dtype=np.complex128)
    gr = np.zeros(len(wr), dtype=np.complex128)
    wm, wr = _f64(wm), _f64(wr)
    lib().orc_add_to_lanczos_gf(norm2, zeta, ei, _dp(a), _dp(b), a.size, isign, _dp(wm), wm.size,
                                gm.ctypes.data_as(c_dp), _dp(wr), wr.size, eps, gr.ctypes.data_as(c_dp))
    return gm, gr


def add_to_lanczos_chi(norm2, ei, beta, alanc, blanc, vm, tau, vr, eps, zeta=1.0):
    """add_to_lanczos_spinChi / _densChi at T = 0: (chi_iv[0..Lmats], chi_tau[0..Ltau], chi_w[Lreal])."""
    a, b = _f64(alanc), _f64(blanc)
    vm, tau, vr = _f64(vm), _f64(tau), _f64(vr)
    civ = np.zeros(len(vm))
    ctau = np.zeros(len(tau))
    cw = np.zeros(len(vr), dtype=np.complex128)
    lib().orc_add_to_lanczos_chi(C.c_double(norm2), C.c_double(zeta), C.c_double(ei), C.c_double(beta), _dp(a), _dp(b), a.size,
                                 _dp(vm), len(vm) - 1, _dp(civ), _dp(tau), len(tau) - 1, _dp(ctau),
                                 _dp(vr), len(vr), C.c_double(eps), cw.ctypes.data_as(c_dp))
    return civ, ctau, cw


def tql2(d, e):
    """Eigen-decomposition of the symmetric tridiagonal (d, e[1:]); returns (w, Z)."""
    n = len(d)
    dd = _f64(d).copy()
    ee = np.zeros(n + 1)
    ee[1:n] = np.asarray(e)[1:n]
    z = np.eye(n).ravel(order="F").copy()
    rc = lib().orc_tql2(n, _dp(dd), _dp(ee), _dp(z))
    assert rc == 0
    return dd, z.reshape((n, n), order="F")


def spmatvec_mpi_prebuilt(sectors, v, hv, nthreads):
    """spMatVec_MPI_main (ED_HAMILTONIAN_SPARSE_HxV.f90:568-694) for all ranks of a split at once, INCLUDING the two
    vector_transpose_MPI exchanges, on pre-built per-rank sectors; rank r runs on thread r % nthreads."""
    n = len(sectors)
    arr = (C.POINTER(_Sector) * n)(*[s.p for s in sectors])
    lib().orc_spmatvec_mpi_main_prebuilt(arr, n, nthreads, _dp(v), _dp(hv))


def spmatvec_blocks_mt(sectors, v_full, outs, nthreads):
    """Time-able multi-threaded sample: one column block (sector view) per thread."""
    n = len(sectors)
    arr = (C.POINTER(_Sector) * n)(*[s.p for s in sectors])
    po = (c_dp * n)(*[_dp(o) for o in outs])
    v = _f64(v_full)
    return lib().orc_spmatvec_blocks_mt(arr, n, nthreads, _dp(v), po)
