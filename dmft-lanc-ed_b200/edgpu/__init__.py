"""Host-side mirror of the reference's operator interface over the C-ABI library libedgpu.so.

Names follow the reference (ED_HAMILTONIAN.f90, ED_DIAG.f90, ED_GF_NORMAL.f90):
``build_Hv_sector`` / ``delete_Hv_sector`` / ``vecDim_Hv_sector``, the ``spHtimesV_p`` operator,
``sp_lanc_eigh`` / ``sp_lanc_tridiag``.  Everything goes through ``include/edgpu.h`` with plain
pointers (ctypes); there is no CPU fallback -- if the CUDA library or a B200 is missing every
compute call raises.
"""
import ctypes as C
import os

import numpy as np

from . import configs  # noqa: F401

_PKG = os.path.dirname(os.path.abspath(__file__))
# EDGPU_LIB: an alternative build of the same library (kernel-variant experiments, tools/variants.sh)
_LIBPATH = os.environ.get("EDGPU_LIB") or os.path.join(os.path.dirname(_PKG), "libedgpu.so")
_LIB = None
_TESTLIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)

ALGO_AUTO, ALGO_GATHER, ALGO_TILED, ALGO_FAST = 0, 1, 2, 3

# every symbol include/edgpu.h declares (checked by tests/test_abi.py against the header)
ABI_SYMBOLS = [
    "edgpu_create", "edgpu_destroy", "edgpu_set_params", "edgpu_last_error", "edgpu_set_option",
    "edgpu_device_count", "edgpu_comm_unique_id", "edgpu_comm_init", "edgpu_comm_finalize",
    "edgpu_get_sector", "edgpu_get_nup_ndw", "edgpu_split", "edgpu_transpose_plan", "edgpu_build_hv_sector",
    "edgpu_delete_hv_sector", "edgpu_vecdim_hv_sector", "edgpu_hxv", "edgpu_sphtimesv",
    "edgpu_hxv_device", "edgpu_sp_lanc_eigh", "edgpu_sp_lanc_tridiag", "edgpu_diag_sectors", "edgpu_gf_set_state", "edgpu_gf_set_state_from_eigh",
    "edgpu_gf_chains", "edgpu_add_to_lanczos_gf", "edgpu_chi_chains", "edgpu_add_to_lanczos_chi", "edgpu_get_sector_orbs", "edgpu_get_qn_orbs", "edgpu_get_orbs_dims",
    "edgpu_get_orbs_factor", "edgpu_get_orbs_diag", "edgpu_observables_normal", "edgpu_get_dims", "edgpu_get_sector_map",
    "edgpu_get_csr", "edgpu_get_diag", "edgpu_dev_alloc", "edgpu_dev_free", "edgpu_dev_upload",
    "edgpu_dev_download", "edgpu_dev_fill_bench_vector", "edgpu_sync", "edgpu_dev_dot", "edgpu_time_hxv_device",
    "edgpu_time_lanczos_device", "edgpu_launch_count", "edgpu_time_hxv_passes", "edgpu_halo_info",
]


class EdgpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("edgpu error %d: %s" % (code, msg))
        self.code = code


class Observables(C.Structure):
    """edgpu_observables (include/edgpu.h): ED_OBSERVABLES.f90 outputs of one state."""
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


class Params(C.Structure):
    _fields_ = [("norb", C.c_int32), ("nbath", C.c_int32), ("nspin", C.c_int32), ("hfmode", C.c_int32),
                ("ed_sparse_h", C.c_int32), ("nph", C.c_int32), ("ed_total_ud", C.c_int32),
                ("bath_type", C.c_int32), ("uloc", C.c_double * 5), ("ust", C.c_double), ("jh", C.c_double),
                ("jx", C.c_double), ("jp", C.c_double), ("xmu", C.c_double),
                ("imphloc", c_dp), ("bath_e", c_dp), ("bath_v", c_dp), ("g_ph", C.c_double * 5), ("w0_ph", C.c_double),
                ("bath_h", c_dp)]


def lib():
    """Load libedgpu.so (built in-tree by __graft_entry__.build()); raises if it is missing."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(_LIBPATH):
            raise ImportError("%s not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; "
                              "g.build()'); there is no CPU fallback" % _LIBPATH)
        L = C.CDLL(_LIBPATH, mode=C.RTLD_GLOBAL)
        L.edgpu_last_error.restype = C.c_char_p
        L.edgpu_create.argtypes = [C.POINTER(Params), C.c_int, C.POINTER(C.c_void_p)]
        L.edgpu_destroy.argtypes = [C.c_void_p]
        L.edgpu_set_params.argtypes = [C.c_void_p, C.POINTER(Params)]
        L.edgpu_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
        L.edgpu_device_count.argtypes = [c_ip]
        L.edgpu_comm_unique_id.argtypes = [C.c_char_p]
        L.edgpu_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p]
        L.edgpu_comm_finalize.argtypes = [C.c_void_p]
        L.edgpu_get_sector.argtypes = [C.c_void_p, C.c_int, C.c_int, c_ip]
        L.edgpu_get_nup_ndw.argtypes = [C.c_void_p, C.c_int, c_ip, c_ip]
        L.edgpu_split.argtypes = [C.c_int64, C.c_int, C.c_int, c_i64p, c_i64p]
        L.edgpu_split.restype = None
        L.edgpu_transpose_plan.argtypes = [C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, c_i64p, c_i64p, c_i64p, c_i64p]
        L.edgpu_transpose_plan.restype = None
        L.edgpu_build_hv_sector.argtypes = [C.c_void_p, C.c_int]
        L.edgpu_delete_hv_sector.argtypes = [C.c_void_p]
        L.edgpu_vecdim_hv_sector.argtypes = [C.c_void_p, C.c_int, c_i64p]
        L.edgpu_hxv.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.edgpu_sphtimesv.argtypes = [c_i32p, C.c_void_p, C.c_void_p]
        L.edgpu_sphtimesv.restype = None
        L.edgpu_hxv_device.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.edgpu_sp_lanc_eigh.argtypes = [C.c_void_p, c_dp, c_dp, C.c_int64, C.c_int, C.c_int, C.c_double, C.c_int,
                                         c_ip, c_dp, c_dp]
        L.edgpu_sp_lanc_tridiag.argtypes = [C.c_void_p, c_dp, C.c_int64, c_dp, c_dp, C.c_int, C.c_double]
        L.edgpu_diag_sectors.argtypes = [C.c_void_p, C.c_int, c_ip, C.c_int, C.c_double, C.c_int, C.c_int, c_dp, c_ip, c_ip]
        L.edgpu_gf_set_state.argtypes = [C.c_void_p, C.c_int, c_dp, C.c_int64, C.c_double]
        L.edgpu_gf_set_state_from_eigh.argtypes = [C.c_void_p]
        L.edgpu_gf_chains.argtypes = [C.c_void_p, C.c_int, c_ip, c_ip, c_ip, C.c_int, C.c_double, c_dp, c_ip, c_dp, c_dp]
        L.edgpu_get_sector_orbs.argtypes = [C.c_void_p, c_ip, c_ip, c_ip]
        L.edgpu_get_qn_orbs.argtypes = [C.c_void_p, C.c_int, c_ip, c_ip]
        L.edgpu_get_orbs_dims.argtypes = [C.c_void_p, c_i64p, c_i64p]
        L.edgpu_get_orbs_factor.argtypes = [C.c_void_p, C.c_int, c_i32p, c_i64p, c_i64p, c_i64p, c_dp]
        L.edgpu_get_orbs_diag.argtypes = [C.c_void_p, c_dp]
        L.edgpu_chi_chains.argtypes = [C.c_void_p, C.c_int, C.c_int, c_ip, c_ip, C.c_int, C.c_double, c_dp, c_ip, c_dp, c_dp]
        L.edgpu_add_to_lanczos_chi.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, c_dp, c_dp, C.c_int,
                                               c_dp, C.c_int, c_dp, c_dp, C.c_int, c_dp, c_dp, C.c_int, C.c_double, c_dp]
        L.edgpu_add_to_lanczos_gf.argtypes = [C.c_double, C.c_double, C.c_double, c_dp, c_dp, C.c_int, C.c_int,
                                              c_dp, C.c_int, c_dp]
        L.edgpu_observables_normal.argtypes = [C.c_void_p, C.c_double, C.POINTER(Observables)]
        L.edgpu_get_dims.argtypes = [C.c_void_p, c_i64p, c_i64p, c_i64p, c_i64p, c_i64p]
        L.edgpu_get_sector_map.argtypes = [C.c_void_p, C.c_int, c_i32p]
        L.edgpu_get_csr.argtypes = [C.c_void_p, C.c_int, c_i64p, c_i64p, c_i64p, c_i64p, c_dp]
        L.edgpu_get_diag.argtypes = [C.c_void_p, c_dp, C.c_int64]
        L.edgpu_dev_alloc.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p)]
        L.edgpu_dev_free.argtypes = [C.c_void_p, C.c_void_p]
        L.edgpu_dev_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.edgpu_dev_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.edgpu_dev_fill_bench_vector.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
        L.edgpu_sync.argtypes = [C.c_void_p]
        L.edgpu_dev_dot.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, c_dp]
        L.edgpu_time_hxv_device.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, c_dp]
        L.edgpu_time_lanczos_device.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, c_dp]
        L.edgpu_launch_count.argtypes = [C.c_void_p, c_i64p]
        L.edgpu_halo_info.argtypes = [C.c_void_p, c_i64p, c_i64p, c_ip]
        L.edgpu_time_hxv_passes.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, c_ip, c_dp, C.c_char_p]
        _LIB = L
    return _LIB


def selftest_lib():
    """libedgpu_selftest.so: test instrumentation only (include/edgpu_selftest.h), a separate library linked
    against libedgpu.so.  Nothing in the product path loads it."""
    global _TESTLIB
    if _TESTLIB is None:
        lib()
        path = os.path.join(os.path.dirname(_LIBPATH), "libedgpu_selftest.so")
        if not os.path.exists(path):
            raise ImportError("%s not found: build it with make -C dmft-lanc-ed_b200/csrc" % path)
        _TESTLIB = C.CDLL(path, mode=C.RTLD_GLOBAL)
    return _TESTLIB


def _ck(rc):
    if rc != 0:
        raise EdgpuError(rc, lib().edgpu_last_error().decode())


def _dp(a):
    return a.ctypes.data_as(c_dp)


def device_count():
    n = C.c_int(0)
    try:
        _ck(lib().edgpu_device_count(C.byref(n)))
    except EdgpuError:
        return 0
    return n.value


def split(n, nranks, rank):
    """(q, off) of the reference's +1-for-the-first-mod split (ED_HAMILTONIAN.f90:96-110)."""
    q, off = C.c_int64(0), C.c_int64(0)
    lib().edgpu_split(n, nranks, rank, C.byref(q), C.byref(off))
    return q.value, off.value


def transpose_plan(dimup, dimdw, nranks, rank, direction):
    """(soff, scnt, roff, rcnt) of the grouped all-to-all transpose, as comm.cu uses them."""
    arrs = [np.zeros(nranks, dtype=np.int64) for _ in range(4)]
    lib().edgpu_transpose_plan(dimup, dimdw, nranks, rank, direction, *[a.ctypes.data_as(c_i64p) for a in arrs])
    return arrs


def comm_unique_id():
    buf = C.create_string_buffer(128)
    _ck(lib().edgpu_comm_unique_id(buf))
    return buf.raw


class Solver:
    """Module-global ED state + the live sector (ED_VARS_GLOBAL / ED_HAMILTONIAN_COMMON)."""

    def __init__(self, norb, nbath, nspin=1, uloc=(2.0,), ust=0.0, jh=0.0, jx=0.0, jp=0.0, xmu=0.0, hfmode=True,
                 imphloc=None, bath_e=None, bath_v=None, ed_sparse_h=True, device=-1, ed_total_ud=True, nph=0, g_ph=(), w0_ph=0.0,
                 bath_type=0, bath_h=None):
        """bath_type 0 normal: bath_e, bath_v (nspin,norb,nbath); 1 hybrid: bath_e (nspin,1,nbath), bath_v (nspin,norb,nbath);
        2 replica: bath_v (nspin,nbath), bath_h (nspin,nspin,norb,norb,nbath) = the replicas' Hbath matrices."""
        self.norb, self.nbath, self.nspin = norb, nbath, nspin
        self.bath_type, self._bath_h = int(bath_type), bath_h
        self.ed_total_ud = bool(ed_total_ud)
        self.nph, self.g_ph, self.w0_ph = int(nph), tuple(g_ph), float(w0_ph)
        self.ns = norb + nbath if self.bath_type == 1 else (nbath + 1) * norb
        if self.bath_type == 0 and (bath_e is None or bath_v is None):
            bath_e, bath_v = configs.init_dmft_bath(norb, nbath, nspin)
        self._keep = self._pack(uloc, ust, jh, jx, jp, xmu, hfmode, imphloc, bath_e, bath_v, ed_sparse_h)
        self.h = C.c_void_p(None)
        _ck(lib().edgpu_create(C.byref(self._keep[0]), device, C.byref(self.h)))
        self.rank, self.nranks = 0, 1
        self.nloc = 0

    def _pack(self, uloc, ust, jh, jx, jp, xmu, hfmode, imphloc, bath_e, bath_v, ed_sparse_h):
        p = Params()
        p.norb, p.nbath, p.nspin = self.norb, self.nbath, self.nspin
        p.hfmode, p.ed_sparse_h, p.nph, p.ed_total_ud = int(bool(hfmode)), int(bool(ed_sparse_h)), int(getattr(self, "nph", 0)), int(getattr(self, "ed_total_ud", True))
        for i, g in enumerate(getattr(self, "g_ph", ())):
            p.g_ph[i] = float(g)
        p.w0_ph = float(getattr(self, "w0_ph", 0.0))
        for i in range(5):
            p.uloc[i] = float(uloc[i]) if i < len(uloc) else 0.0
        p.ust, p.jh, p.jx, p.jp, p.xmu = ust, jh, jx, jp, xmu
        bt = int(getattr(self, "bath_type", 0))
        p.bath_type = bt
        shp = (self.nspin, self.norb, self.nbath)
        if bt == 0:
            fe = np.ravel(np.asarray(bath_e, dtype=np.float64).reshape(shp, order="F"), order="F").copy()
            fv = np.ravel(np.asarray(bath_v, dtype=np.float64).reshape(shp, order="F"), order="F").copy()
        else:
            fe = np.ravel(np.asarray(bath_e, dtype=np.float64), order="F").copy() if bath_e is not None else np.zeros(1)
            fv = np.ravel(np.asarray(bath_v, dtype=np.float64), order="F").copy()
        fb = None
        if bt == 2:
            fb = np.ravel(np.asarray(getattr(self, "_bath_h"), dtype=np.float64), order="F").copy()
            p.bath_h = _dp(fb)
        if imphloc is None:
            imphloc = np.zeros((self.nspin, self.nspin, self.norb, self.norb))
        fh = np.ravel(np.asarray(imphloc, dtype=np.float64), order="F").copy()
        p.imphloc, p.bath_e, p.bath_v = _dp(fh), _dp(fe), _dp(fv)
        return (p, fe, fv, fh, fb)

    def close(self):
        if self.h:
            lib().edgpu_destroy(self.h)
            self.h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key, value):
        _ck(lib().edgpu_set_option(self.h, key.encode(), int(value)))

    # ---- ed_set_MpiComm -----------------------------------------------------------------------
    def set_comm(self, rank, nranks, unique_id):
        _ck(lib().edgpu_comm_init(self.h, rank, nranks, unique_id))
        self.rank, self.nranks = rank, nranks

    # ---- sector bookkeeping ----------------------------------------------------------------------
    def get_sector(self, nup, ndw):
        s = C.c_int(0)
        _ck(lib().edgpu_get_sector(self.h, nup, ndw, C.byref(s)))
        return s.value

    # ed_total_ud = F: one (Nup, Ndw) pair per orbital
    def get_sector_orbs(self, nups, ndws):
        n = self.norb
        s = C.c_int(0)
        _ck(lib().edgpu_get_sector_orbs(self.h, (C.c_int * n)(*nups), (C.c_int * n)(*ndws), C.byref(s)))
        return s.value

    def get_qn_orbs(self, isector):
        n = self.norb
        a, b = (C.c_int * n)(), (C.c_int * n)()
        _ck(lib().edgpu_get_qn_orbs(self.h, isector, a, b))
        return list(a), list(b)

    def orbs_dims(self):
        d = (C.c_int64 * (2 * self.norb))()
        tot = C.c_int64(0)
        _ck(lib().edgpu_get_orbs_dims(self.h, d, C.byref(tot)))
        return list(d), tot.value

    def orbs_factor(self, f):
        """(map, rowptr, cols, vals) of factor f of the live ed_total_ud = F sector."""
        dims, _ = self.orbs_dims()
        nnz = C.c_int64(0)
        m = np.zeros(dims[f], np.int32)
        rp = np.zeros(dims[f] + 1, np.int64)
        i64p = C.POINTER(C.c_int64)
        _ck(lib().edgpu_get_orbs_factor(self.h, f, m.ctypes.data_as(c_i32p), C.byref(nnz), rp.ctypes.data_as(i64p), None, None))
        cols = np.zeros(max(nnz.value, 1), np.int64)
        vals = np.zeros(max(nnz.value, 1))
        _ck(lib().edgpu_get_orbs_factor(self.h, f, None, None, None, cols.ctypes.data_as(i64p), _dp(vals)))
        return m, rp, cols[:nnz.value], vals[:nnz.value]

    def orbs_diag(self):
        d = np.zeros(self.nloc)
        _ck(lib().edgpu_get_orbs_diag(self.h, _dp(d)))
        return d

    def get_nup_ndw(self, isector):
        a, b = C.c_int(0), C.c_int(0)
        _ck(lib().edgpu_get_nup_ndw(self.h, isector, C.byref(a), C.byref(b)))
        return a.value, b.value

    def vecDim_Hv_sector(self, isector):
        v = C.c_int64(0)
        _ck(lib().edgpu_vecdim_hv_sector(self.h, isector, C.byref(v)))
        return v.value

    def build_Hv_sector(self, isector):
        _ck(lib().edgpu_build_hv_sector(self.h, isector))
        d = [C.c_int64(0) for _ in range(5)]
        _ck(lib().edgpu_get_dims(self.h, *[C.byref(x) for x in d]))
        self.dimup, self.dimdw, self.qdw, self.ishift, self.nloc = [x.value for x in d]

    def delete_Hv_sector(self):
        _ck(lib().edgpu_delete_hv_sector(self.h))
        self.nloc = 0

    # ---- the operator ------------------------------------------------------------------------------
    def spHtimesV(self, v, out=None):
        """spHtimesV_p(Nloc, v, Hv) with host arrays."""
        v = np.ascontiguousarray(v, dtype=np.float64)
        hv = np.empty_like(v) if out is None else out
        _ck(lib().edgpu_hxv(self.h, v.size, v.ctypes.data, hv.ctypes.data))
        return hv

    def spHtimesV_ptr(self, nloc, v_ptr, hv_ptr):
        _ck(lib().edgpu_hxv(self.h, nloc, v_ptr, hv_ptr))

    def spHtimesV_fortran(self, v):
        """Through the procedure-pointer compatible symbol edgpu_sphtimesv(Nloc,v,Hv)."""
        v = np.ascontiguousarray(v, dtype=np.float64)
        hv = np.empty_like(v)
        n = C.c_int32(v.size)
        lib().edgpu_sphtimesv(C.byref(n), v.ctypes.data, hv.ctypes.data)
        return hv

    def hxv_device(self, d_v, d_hv):
        _ck(lib().edgpu_hxv_device(self.h, self.nloc, d_v, d_hv))

    # ---- Lanczos -----------------------------------------------------------------------------------------
    def sp_lanc_eigh(self, vect=None, nitermax=512, threshold=1e-18, ncheck=10, iverbose=False):
        """Returns (egs, vect, alanc, blanc) -- ED_DIAG.f90:177-185."""
        nit = int(min(self.dimup * self.dimdw, nitermax))
        v = np.zeros(self.nloc) if vect is None else np.ascontiguousarray(vect, dtype=np.float64).copy()
        egs = C.c_double(0.0)
        nl = C.c_int(0)
        a = np.zeros(nit + 2)
        b = np.zeros(nit + 2)
        _ck(lib().edgpu_sp_lanc_eigh(self.h, C.byref(egs), _dp(v), v.size, nit, int(iverbose), threshold, ncheck,
                                     C.byref(nl), _dp(a), _dp(b)))
        return egs.value, v, a[:nl.value].copy(), b[:nl.value].copy()

    def sp_lanc_eigh_ptr(self, vect_ptr, nitermax=512, threshold=1e-18, ncheck=10):
        """The same call on a caller-owned host buffer given by address (start vector in, eigenvector out; e.g. pinned
        memory): no copies on the Python side.  Returns (egs, nlanc)."""
        nit = int(min(self.dimup * self.dimdw, nitermax))
        egs = C.c_double(0.0)
        nl = C.c_int(0)
        _ck(lib().edgpu_sp_lanc_eigh(self.h, C.byref(egs), C.cast(C.c_void_p(int(vect_ptr)), c_dp), self.nloc, nit, 0, threshold, ncheck,
                                     C.byref(nl), None, None))
        return egs.value, nl.value

    def sp_lanc_tridiag(self, vin, nlanc, threshold=1e-12):
        v = np.ascontiguousarray(vin, dtype=np.float64)
        a = np.zeros(nlanc)
        b = np.zeros(nlanc)
        _ck(lib().edgpu_sp_lanc_tridiag(self.h, _dp(v), v.size, _dp(a), _dp(b), nlanc, threshold))
        return a, b

    def diag_sectors(self, sectors, nitermax=512, threshold=1e-18, ncheck=10, twin=False):
        """ed_diag_d's sector loop (ED_DIAG.f90:83-276): (e0[], nlanc[], best) -- the lowest sector's eigenvector stays
        on the device as the state of the chains / observables."""
        n = len(sectors)
        sec = (C.c_int * n)(*sectors)
        e0 = np.zeros(n)
        nl = (C.c_int * n)()
        best = C.c_int(-1)
        _ck(lib().edgpu_diag_sectors(self.h, n, sec, nitermax, threshold, ncheck, int(bool(twin)), _dp(e0), nl, C.byref(best)))
        return e0, np.array(nl[:]), best.value

    # ---- Green's function chains -----------------------------------------------------------------------
    def gf_set_state(self, isector, gs, e0):
        gs = np.ascontiguousarray(gs, dtype=np.float64)
        _ck(lib().edgpu_gf_set_state(self.h, isector, _dp(gs), gs.size, e0))

    def gf_set_state_from_eigh(self):
        """Keep the eigenvector of the last sp_lanc_eigh (sector still live) on the device as the chains' state."""
        _ck(lib().edgpu_gf_set_state_from_eigh(self.h))

    def gf_chains(self, channels, nlanc_max=200, threshold=1e-12):
        """channels: list of (iorb, ispin, +1|-1).  Returns list of dicts(norm2, nlanc, alanc, blanc)."""
        n = len(channels)
        io = (C.c_int * n)(*[c[0] for c in channels])
        sp = (C.c_int * n)(*[c[1] for c in channels])
        ar = (C.c_int * n)(*[c[2] for c in channels])
        norm2 = np.zeros(n)
        nl = (C.c_int * n)()
        a = np.zeros((n, nlanc_max))
        b = np.zeros((n, nlanc_max))
        _ck(lib().edgpu_gf_chains(self.h, n, io, sp, ar, nlanc_max, threshold, _dp(norm2), nl, _dp(a), _dp(b)))
        return [dict(norm2=norm2[k], nlanc=nl[k], alanc=a[k, :nl[k]].copy(), blanc=b[k, :nl[k]].copy())
                for k in range(n)]

    def chi_chains(self, kind, channels, nlanc_max=200, threshold=1e-12):
        """Susceptibility chains (lanc_ed_build_spinChi_* kind=0 / lanc_ed_build_densChi_* kind=1).  channels: list of
        (iorb, jorb): iorb == jorb one orbital, iorb == 0 total, iorb != jorb mixed.  Same output as gf_chains."""
        n = len(channels)
        io = (C.c_int * n)(*[c[0] for c in channels])
        jo = (C.c_int * n)(*[c[1] for c in channels])
        norm2 = np.zeros(n)
        nl = (C.c_int * n)()
        a = np.zeros((n, nlanc_max))
        b = np.zeros((n, nlanc_max))
        _ck(lib().edgpu_chi_chains(self.h, int(kind), n, io, jo, nlanc_max, threshold, _dp(norm2), nl, _dp(a), _dp(b)))
        return [dict(norm2=norm2[k], nlanc=nl[k], alanc=a[k, :nl[k]].copy(), blanc=b[k, :nl[k]].copy())
                for k in range(n)]

    def observables(self, zeta=1.0):
        """lanc_observables + lanc_local_energy of the state kept for the chains (collective when sharded)."""
        o = Observables()
        _ck(lib().edgpu_observables_normal(self.h, zeta, C.byref(o)))
        return o.as_dict()

    # ---- introspection ---------------------------------------------------------------------------------------
    def sector_map(self, which):
        n = self.dimdw if which else self.dimup
        m = np.zeros(n, dtype=np.int32)
        _ck(lib().edgpu_get_sector_map(self.h, which, m.ctypes.data_as(c_i32p)))
        return m

    def csr(self, which):
        nrow, nnz = C.c_int64(0), C.c_int64(0)
        _ck(lib().edgpu_get_csr(self.h, which, C.byref(nrow), C.byref(nnz), None, None, None))
        rp = np.zeros(nrow.value + 1, dtype=np.int64)
        cols = np.zeros(max(nnz.value, 1), dtype=np.int64)
        vals = np.zeros(max(nnz.value, 1))
        if nrow.value:
            _ck(lib().edgpu_get_csr(self.h, which, C.byref(nrow), C.byref(nnz), rp.ctypes.data_as(c_i64p),
                                    cols.ctypes.data_as(c_i64p), _dp(vals)))
        return rp, cols[:nnz.value], vals[:nnz.value]

    def diag(self):
        """spH0d of the local electron part (DimUp * mpiQdw values; the phonon slabs share it)."""
        d = np.zeros(self.dimup * self.qdw)
        _ck(lib().edgpu_get_diag(self.h, _dp(d), d.size))
        return d

    # ---- device-resident helpers ----------------------------------------------------------------------------
    def dev_alloc(self, nbytes):
        p = C.c_void_p(None)
        _ck(lib().edgpu_dev_alloc(self.h, nbytes, C.byref(p)))
        return p

    def dev_free(self, p):
        _ck(lib().edgpu_dev_free(self.h, p))

    def dev_upload(self, p, arr):
        arr = np.ascontiguousarray(arr)
        _ck(lib().edgpu_dev_upload(self.h, p, arr.ctypes.data, arr.nbytes))

    def dev_download(self, p, arr):
        _ck(lib().edgpu_dev_download(self.h, arr.ctypes.data, p, arr.nbytes))

    def dev_fill_bench_vector(self, p, nloc, global_offset):
        _ck(lib().edgpu_dev_fill_bench_vector(self.h, p, nloc, global_offset))

    def sync(self):
        _ck(lib().edgpu_sync(self.h))

    def dev_dot(self, d_a, d_b):
        """<a, b> over the whole sector vector (all ranks): collective."""
        out = C.c_double(0.0)
        _ck(lib().edgpu_dev_dot(self.h, self.nloc, d_a, d_b, C.byref(out)))
        return out.value

    def dev_download_slice(self, p, offset_elems, arr):
        """arr <- device doubles [offset, offset + arr.size) of buffer p."""
        q = C.c_void_p(p.value + 8 * int(offset_elems))
        _ck(lib().edgpu_dev_download(self.h, arr.ctypes.data, q, arr.nbytes))

    def time_hxv_device(self, d_v, d_hv, reps):
        ms = C.c_double(0.0)
        _ck(lib().edgpu_time_hxv_device(self.h, self.nloc, d_v, d_hv, reps, C.byref(ms)))
        return ms.value

    def time_hxv_passes(self, d_v, d_hv, reps):
        """[(kernel name, ms per launch)] of one device-resident H*v, CUDA events between the passes."""
        n = C.c_int(0)
        ms = (C.c_double * 6)()
        names = C.create_string_buffer(192)
        _ck(lib().edgpu_time_hxv_passes(self.h, self.nloc, d_v, d_hv, int(reps), C.byref(n), ms, names))
        return [(names.raw[32 * i:32 * i + 32].split(b"\0")[0].decode(), ms[i] / reps) for i in range(n.value)]

    def time_lanczos_device(self, d_v0, reps):
        ms = C.c_double(0.0)
        _ck(lib().edgpu_time_lanczos_device(self.h, self.nloc, d_v0, reps, C.byref(ms)))
        return ms.value

    def halo_info(self):
        """(bytes stored into the peers' halo buffers per H*v, bytes received, column windows) of the live sector."""
        bo, bi, w = C.c_int64(0), C.c_int64(0), C.c_int(0)
        _ck(lib().edgpu_halo_info(self.h, C.byref(bo), C.byref(bi), C.byref(w)))
        return bo.value, bi.value, w.value

    def launch_count(self):
        n = C.c_int64(0)
        _ck(lib().edgpu_launch_count(self.h, C.byref(n)))
        return n.value


def add_to_lanczos_chi(norm2, ei, beta, alanc, blanc, vm, tau, vr, eps, zeta=1.0):
    """add_to_lanczos_spinChi / _densChi at T = 0: (chi_iv[0..Lmats], chi_tau[0..Ltau], chi_w[Lreal])."""
    a = np.ascontiguousarray(alanc, dtype=np.float64)
    b = np.ascontiguousarray(blanc, dtype=np.float64)
    vm = np.ascontiguousarray(vm, dtype=np.float64)
    tau = np.ascontiguousarray(tau, dtype=np.float64)
    vr = np.ascontiguousarray(vr, dtype=np.float64)
    civ, ctau, cw = np.zeros(len(vm)), np.zeros(len(tau)), np.zeros(len(vr), dtype=np.complex128)
    _ck(lib().edgpu_add_to_lanczos_chi(norm2, zeta, ei, beta, _dp(a), _dp(b), a.size, _dp(vm), len(vm) - 1, _dp(civ),
                                       _dp(tau), len(tau) - 1, _dp(ctau), _dp(vr), len(vr), eps, cw.ctypes.data_as(c_dp)))
    return civ, ctau, cw


def add_to_lanczos_gf(norm2, ei, alanc, blanc, isign, z, zeta=1.0):
    """G(z) contribution of one chain (ED_GF_NORMAL.f90:599-654, T=0)."""
    a = np.ascontiguousarray(alanc, dtype=np.float64)
    b = np.ascontiguousarray(blanc, dtype=np.float64)
    z = np.ascontiguousarray(z, dtype=np.complex128)
    g = np.zeros_like(z)
    _ck(lib().edgpu_add_to_lanczos_gf(norm2, zeta, ei, _dp(a), _dp(b), a.size, isign,
                                      z.ctypes.data_as(c_dp), z.size, g.ctypes.data_as(c_dp)))
    return g


def sigma_normal(z, g, xmu, hloc, bath_e, bath_v):
    """build_sigma_normal for bath_type normal, one (ispin,iorb): Sigma = G0^-1 - G^-1 with
    G0^-1 = z + xmu - impHloc - sum_k V_k^2/(z - e_k) (ED_GF_NORMAL.f90:935-1002,
    ED_BATH_FUNCTIONS.f90:43-77,163-195).  O(L*Nbath) host arithmetic on the chain output."""
    z = np.asarray(z, dtype=np.complex128)
    delta = np.zeros_like(z)
    for e, v in zip(np.ravel(bath_e), np.ravel(bath_v)):
        delta = delta + v * v / (z - e)
    invg0 = z + xmu - hloc - delta
    return invg0 - 1.0 / np.asarray(g), invg0
