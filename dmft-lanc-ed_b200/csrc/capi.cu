// capi.cu -- context, sector lifecycle (build_Hv_sector / delete_Hv_sector), basis and stored
// factor construction on device, introspection.  See include/edgpu.h for the reference
// interfaces each entry point replaces.
#include <cub/device/device_scan.cuh>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "engine.h"

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
static edgpu_ctx *g_current = nullptr;   // context of the last build_hv_sector (spHtimesV_p analogue)

int edgpu_set_err(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
extern "C" const char *edgpu_last_error(void) { return g_err; }

extern "C" int edgpu_device_count(int *n) {
  int k = 0;
  cudaError_t e = cudaGetDeviceCount(&k);
  if (e != cudaSuccess) { *n = 0; return edgpu_set_err(EDGPU_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
  *n = k;
  return EDGPU_OK;
}

// ------------------------------------------------------------------------------------------
// parameters
// ------------------------------------------------------------------------------------------
static int load_params(edgpu_ctx *c, const edgpu_params *p) {
  if (!p) return edgpu_set_err(EDGPU_ERR_INVALID, "params == NULL");
  if (p->norb < 1 || p->norb > EDGPU_MAX_ORB) return edgpu_set_err(EDGPU_ERR_INVALID, "NORB out of range");
  if (p->nspin < 1 || p->nspin > 2) return edgpu_set_err(EDGPU_ERR_INVALID, "NSPIN out of range");
  if (p->nph < 0 || p->nph > 1000) return edgpu_set_err(EDGPU_ERR_INVALID, "NPH out of range");
  if (p->nph > 0 && !p->ed_total_ud) return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "phonons with ed_total_ud = F are not built");
  if (!p->ed_total_ud && p->norb > 1 && (p->jx != 0.0 || p->jp != 0.0))     // ED_SETUP.f90:69-71
    return edgpu_set_err(EDGPU_ERR_INVALID, "ed_total_ud = F can not be used with Jx != 0 or Jp != 0");
  const int bt = p->bath_type;
  if (bt < 0 || bt > 2) return edgpu_set_err(EDGPU_ERR_INVALID, "BATH_TYPE must be 0 (normal), 1 (hybrid) or 2 (replica)");
  if (!p->ed_total_ud && bt != 0) return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "ed_total_ud = F is built for the normal bath only (hybrid: ED_SETUP.f90:70)");
  int ns = (bt == 1) ? p->norb + p->nbath : (p->nbath + 1) * p->norb;   // ED_SETUP.f90:113-121
  if (p->nbath < 1 || ns > EDGPU_MAX_SITES - 1) return edgpu_set_err(EDGPU_ERR_INVALID, "Ns = %d out of range", ns);
  if (!p->bath_v || (bt != 2 && !p->bath_e) || (bt == 2 && !p->bath_h)) return edgpu_set_err(EDGPU_ERR_INVALID, "bath arrays == NULL");
  if (bt == 2 && (p->norb > 3 || p->nbath > EDGPU_MAX_REPL_BATH))
    return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "replica bath: Norb <= 3 and Nbath <= %d", EDGPU_MAX_REPL_BATH);
  c->hp = *p;
  c->ns = ns;
  c->dimph = p->nph + 1;
  const int nsn = p->nspin, sl = p->nspin - 1;
  size_t nh = (size_t)nsn * nsn * p->norb * p->norb;
  const size_t ne = (bt == 0) ? (size_t)nsn * p->norb * p->nbath : (bt == 1 ? (size_t)nsn * p->nbath : 0);
  const size_t nv = (bt == 2) ? (size_t)nsn * p->nbath : (size_t)nsn * p->norb * p->nbath;
  const size_t nbh = (bt == 2) ? nh * p->nbath : 0;
  c->h_hloc.assign(nh, 0.0);
  if (p->imphloc) memcpy(c->h_hloc.data(), p->imphloc, nh * sizeof(double));
  c->h_be.assign(std::max<size_t>(ne, 1), 0.0);
  if (ne) memcpy(c->h_be.data(), p->bath_e, ne * sizeof(double));
  c->h_bv.assign(p->bath_v, p->bath_v + nv);
  c->h_bh.assign(std::max<size_t>(nbh, 1), 0.0);
  if (nbh) memcpy(c->h_bh.data(), p->bath_h, nbh * sizeof(double));
  c->hp.imphloc = c->h_hloc.data();
  c->hp.bath_e = c->h_be.data();
  c->hp.bath_v = c->h_bv.data();
  c->hp.bath_h = nbh ? c->h_bh.data() : nullptr;
  DevParams &d = c->dp;
  memset(&d, 0, sizeof(d));
  d.norb = p->norb; d.nbath = p->nbath; d.ns = ns; d.hfmode = p->hfmode; d.nspin = p->nspin;
  d.bath_type = bt; d.nfoo = (bt == 1) ? 1 : p->norb;
  d.jhflag = (p->norb > 1 && (p->jx != 0.0 || p->jp != 0.0));   // ED_SETUP.f90:147-148
  for (int i = 0; i < EDGPU_MAX_ORB; i++) d.uloc[i] = (i < p->norb) ? p->uloc[i] : 0.0;
  d.ust = p->ust; d.jh = p->jh; d.jx = p->jx; d.jp = p->jp; d.xmu = p->xmu;
  for (int io = 0; io < p->norb; io++)
    for (int jo = 0; jo < p->norb; jo++) {
      d.hloc_up[io * EDGPU_MAX_ORB + jo] = c->h_hloc[0 + nsn * (0 + nsn * (io + p->norb * jo))];
      d.hloc_dw[io * EDGPU_MAX_ORB + jo] = c->h_hloc[sl + nsn * (sl + nsn * (io + p->norb * jo))];
    }
  // diag_hybr / bath_diag as ed_buildh_main assembles them per bath type (ED_HAMILTONIAN_SPARSE_HxV.f90:46-76)
  auto hb = [&](int is, int io, int jo, int kp) { return c->h_bh[(size_t)is + nsn * ((size_t)is + nsn * ((size_t)io + p->norb * ((size_t)jo + p->norb * (size_t)kp)))]; };
  for (int io = 0; io < p->norb; io++)
    for (int kp = 0; kp < p->nbath; kp++) {
      const int at = io * p->nbath + kp;
      if (bt == 0) {
        d.be_up[at] = c->h_be[0 + nsn * (io + p->norb * kp)];
        d.be_dw[at] = c->h_be[sl + nsn * (io + p->norb * kp)];
        d.bv_up[at] = c->h_bv[0 + nsn * (io + p->norb * kp)];
        d.bv_dw[at] = c->h_bv[sl + nsn * (io + p->norb * kp)];
      } else if (bt == 1) {
        if (io == 0) { d.be_up[at] = c->h_be[0 + nsn * kp]; d.be_dw[at] = c->h_be[sl + nsn * kp]; }
        d.bv_up[at] = c->h_bv[0 + nsn * (io + p->norb * kp)];
        d.bv_dw[at] = c->h_bv[sl + nsn * (io + p->norb * kp)];
      } else {
        d.be_up[at] = hb(0, io, io, kp); d.be_dw[at] = hb(sl, io, io, kp);
        d.bv_up[at] = c->h_bv[0 + nsn * kp]; d.bv_dw[at] = c->h_bv[sl + nsn * kp];
        for (int jo = 0; jo < p->norb; jo++) { d.hb_up[kp * 9 + io * 3 + jo] = hb(0, io, jo, kp); d.hb_dw[kp * 9 + io * 3 + jo] = hb(sl, io, jo, kp); }
      }
    }
  return EDGPU_OK;
}

static int create_device_state(edgpu_ctx *c) {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, c->device));
  if (prop.major < 10) return edgpu_set_err(EDGPU_ERR_NO_DEVICE, "device %s is sm_%d%d; this engine is built for sm_100a only", prop.name, prop.major, prop.minor);
  c->sm_count = prop.multiProcessorCount;
  CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
  CK(cudaEventCreate(&c->ev0));
  CK(cudaEventCreate(&c->ev1));
  CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  // exact binomials by Pascal's rule (== binomial(), ED_SETUP.f90:1017-1035, for these sizes)
  memset(c->h_binom, 0, sizeof(c->h_binom));
  for (int n = 0; n < EDGPU_BINOM_LD; n++) {
    c->h_binom[n * EDGPU_BINOM_LD + 0] = 1;
    for (int k = 1; k <= n; k++) {
      uint64_t v = (uint64_t)c->h_binom[(n - 1) * EDGPU_BINOM_LD + k - 1] +
                   (uint64_t)(k <= n - 1 ? c->h_binom[(n - 1) * EDGPU_BINOM_LD + k] : 0);
      c->h_binom[n * EDGPU_BINOM_LD + k] = (v > 0xffffffffull) ? 0xffffffffu : (uint32_t)v;
    }
  }
  CK(cudaMalloc(&c->d_binom, sizeof(c->h_binom)));
  CK(cudaMemcpy(c->d_binom, c->h_binom, sizeof(c->h_binom), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&c->d_partials, 4096 * sizeof(double)));
  CK(cudaMemset(c->d_partials, 0, 4096 * sizeof(double)));
  CK(cudaMalloc(&c->d_st, sizeof(LancState)));
  CK(cudaMemset(c->d_st, 0, sizeof(LancState)));
  CK(cudaMallocHost(&c->h_pinned, 4096 * sizeof(double)));
  return EDGPU_OK;
}

extern "C" int edgpu_create(const edgpu_params *p, int device, edgpu_ctx **out) {
  if (!out) return edgpu_set_err(EDGPU_ERR_INVALID, "out == NULL");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return edgpu_set_err(EDGPU_ERR_NO_DEVICE, "no CUDA device: this engine has no CPU fallback");
  edgpu_ctx *c = new edgpu_ctx();
  int rc = load_params(c, p);
  if (rc) { delete c; return rc; }
  if (device < 0) cudaGetDevice(&device);
  c->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete c; return edgpu_set_err(EDGPU_ERR_NO_DEVICE, "cudaSetDevice(%d) failed", device); }
  rc = create_device_state(c);
  if (rc) { edgpu_destroy(c); return rc; }                       // one cleanup path: destroy frees whatever exists
  *out = c;
  return EDGPU_OK;
}

extern "C" int edgpu_set_params(edgpu_ctx *c, const edgpu_params *p) {
  if (!c) return edgpu_set_err(EDGPU_ERR_INVALID, "ctx == NULL");
  if (c->hstatus) return edgpu_set_err(EDGPU_ERR_INVALID, "set_params while a sector is live");
  return load_params(c, p);
}

extern "C" int edgpu_set_option(edgpu_ctx *c, const char *key, int64_t value) {
  if (!c || !key) return edgpu_set_err(EDGPU_ERR_INVALID, "bad option call");
  if (!strcmp(key, "hxv_algo")) { c->algo = (int)value; return EDGPU_OK; }
  if (!strcmp(key, "tile_rows")) { c->opt_tile_rows = value; return EDGPU_OK; }
  if (!strcmp(key, "tile_h")) { c->opt_tile_h = value; return EDGPU_OK; }
  if (!strcmp(key, "col_h")) { c->opt_col_h = value; return EDGPU_OK; }
  if (!strcmp(key, "srow_lr")) { c->opt_srow_lr = value; return EDGPU_OK; }
  if (!strcmp(key, "srow_t")) { c->opt_srow_t = value; return EDGPU_OK; }      // t + 1 forces chunks of 2^t low groups
  if (!strcmp(key, "no_fuse")) { c->opt_no_fuse = value; return EDGPU_OK; }    // Lanczos update as a separate pass
  if (!strcmp(key, "halo_ctas")) { c->opt_halo_ctas = value; return EDGPU_OK; }
  if (!strcmp(key, "halo_windows")) { c->opt_halo_windows = value; return EDGPU_OK; }   // before build_Hv_sector
  if (!strcmp(key, "no_overlap")) { c->opt_no_overlap = value; return EDGPU_OK; }
  if (!strcmp(key, "no_batch")) { c->opt_no_batch = value; return EDGPU_OK; }        // GF chains of a sector one after another
  if (!strcmp(key, "no_peer")) { c->opt_no_peer = value; return EDGPU_OK; }
  if (!strcmp(key, "col_cluster")) { c->opt_col_cluster = value; return EDGPU_OK; }
  if (!strcmp(key, "no_uniform")) { c->opt_no_uniform = value; return EDGPU_OK; }
  return edgpu_set_err(EDGPU_ERR_INVALID, "unknown option %s", key);
}

// ------------------------------------------------------------------------------------------
// sector numbering and shard geometry (host arithmetic)
// ------------------------------------------------------------------------------------------
extern "C" int edgpu_get_sector(const edgpu_ctx *c, int nup, int ndw, int *isector) {
  if (!c || nup < 0 || ndw < 0 || nup > c->ns || ndw > c->ns) return edgpu_set_err(EDGPU_ERR_INVALID, "bad (nup,ndw)");
  *isector = 1 + ndw + nup * (c->ns + 1);                   // get_Sector, ED_SETUP.f90:446-457
  return EDGPU_OK;
}
extern "C" int edgpu_get_nup_ndw(const edgpu_ctx *c, int isector, int *nup, int *ndw) {
  if (!c) return edgpu_set_err(EDGPU_ERR_INVALID, "ctx == NULL");
  int nsec = (c->ns + 1) * (c->ns + 1);                     // Nsectors, ED_SETUP.f90:134
  if (isector < 1 || isector > nsec) return edgpu_set_err(EDGPU_ERR_INVALID, "isector out of range");
  int count = isector - 1;                                  // get_Nup/get_Ndw, ED_SETUP.f90:477-500
  *ndw = count % (c->ns + 1);
  *nup = count / (c->ns + 1);
  return EDGPU_OK;
}
extern "C" void edgpu_split(int64_t n, int nranks, int rank, int64_t *q, int64_t *off) {
  int64_t qq = n / nranks, m = n % nranks;                  // ED_HAMILTONIAN.f90:96-110
  if (q) *q = qq + (rank < m ? 1 : 0);
  if (off) *off = (rank < m) ? rank * (qq + 1) : m * (qq + 1) + (rank - m) * qq;
}
static int64_t binom64(const edgpu_ctx *c, int n, int k) {
  if (k < 0 || k > n) return 0;
  return c->h_binom[n * EDGPU_BINOM_LD + k];
}
extern "C" int edgpu_vecdim_hv_sector(const edgpu_ctx *c, int isector, int64_t *vecdim) {
  if (c && !c->hp.ed_total_ud) {                              // product of the 2*Norb word dimensions, single rank
    int nups[EDGPU_MAX_ORB], ndws[EDGPU_MAX_ORB];
    TRY(edgpu_get_qn_orbs(c, isector, nups, ndws));
    int64_t d = 1;
    for (int k = 0; k < c->dp.norb; k++) d *= binom64(c, c->dp.nbath + 1, nups[k]) * binom64(c, c->dp.nbath + 1, ndws[k]);
    *vecdim = d;
    return EDGPU_OK;
  }
  int nup, ndw;
  TRY(edgpu_get_nup_ndw(c, isector, &nup, &ndw));
  int64_t q;
  edgpu_split(binom64(c, c->ns, ndw), c->nranks, c->rank, &q, nullptr);
  *vecdim = binom64(c, c->ns, nup) * q * c->dimph;          // DimUp*mpiQdw*DimPh
  return EDGPU_OK;
}

// ------------------------------------------------------------------------------------------
// basis / factor construction kernels (integer work, bit-exact against the reference)
// ------------------------------------------------------------------------------------------
// build_sector, ED_SETUP.f90:764-777: the ascending popcount-filtered scan, done as a parallel
// filter whose output slot is the closed-form rank of the word.
__global__ void k_build_map(int ns, int n, const uint32_t *__restrict__ binom, int32_t *__restrict__ map) {
  const uint64_t top = 1ull << ns;
  for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < top; s += (uint64_t)gridDim.x * blockDim.x)
    if (__popc((uint32_t)s) == n) map[hd_rank((uint32_t)s, binom)] = (int32_t)s;
}

__global__ void k_factor_count(DevParams P, int spin, const int32_t *__restrict__ map, int64_t n,
                               int32_t *__restrict__ counts) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t cols[EDGPU_MAX_ROW_NNZ]; double vals[EDGPU_MAX_ROW_NNZ];
  counts[i] = hd_factor_row(P, spin, map, n, (uint32_t)map[i], cols, vals);
}
__global__ void k_factor_fill(DevParams P, int spin, const int32_t *__restrict__ map, int64_t n,
                              const int32_t *__restrict__ rowptr, int32_t *__restrict__ ocols,
                              double *__restrict__ ovals) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t cols[EDGPU_MAX_ROW_NNZ]; double vals[EDGPU_MAX_ROW_NNZ];
  int m = hd_factor_row(P, spin, map, n, (uint32_t)map[i], cols, vals);
  int32_t p = rowptr[i];
  for (int k = 0; k < m; k++) { ocols[p + k] = cols[k]; ovals[p + k] = vals[k]; }
}
__global__ void k_dfac(DevParams P, int spin, const int32_t *__restrict__ map, int64_t n, double *__restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = hd_diag_factor(P, spin, (uint32_t)map[i]);
}
// spH0d, stored/H_local.f90:1-80: one value per local row, exact reference summation order
__global__ void k_diag_stored(DevParams P, const int32_t *__restrict__ map_up, const int32_t *__restrict__ map_dw,
                              int64_t dimup, int64_t coloff, int64_t qdw, double *__restrict__ diag) {
  for (int64_t jl = blockIdx.y; jl < qdw; jl += gridDim.y) {
    uint32_t mdw = (uint32_t)map_dw[coloff + jl];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < dimup; i += (int64_t)gridDim.x * blockDim.x)
      diag[i + jl * dimup] = hd_diag_element(P, (uint32_t)map_up[i], mdw);
  }
}
// spH0nd rows (stored/H_non_local.f90:4-85); columns are global electron indices
__global__ void k_nd_count(DevParams P, const int32_t *__restrict__ map_up, const int32_t *__restrict__ map_dw,
                           int64_t dimup, int64_t coloff, int64_t nloc, int64_t *__restrict__ counts) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= nloc) return;
  uint32_t cu[2 * EDGPU_MAX_ORB * EDGPU_MAX_ORB], cd[2 * EDGPU_MAX_ORB * EDGPU_MAX_ORB];
  double v[2 * EDGPU_MAX_ORB * EDGPU_MAX_ORB];
  counts[r] = hd_nonlocal_row(P, (uint32_t)map_up[r % dimup], (uint32_t)map_dw[coloff + r / dimup], cu, cd, v);
}
__global__ void k_nd_fill(DevParams P, const int32_t *__restrict__ map_up, const int32_t *__restrict__ map_dw,
                          int64_t dimup, int64_t dimdw, int64_t coloff, int64_t nloc,
                          const int64_t *__restrict__ rowptr, int64_t *__restrict__ ocols, double *__restrict__ ovals) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= nloc) return;
  uint32_t cu[2 * EDGPU_MAX_ORB * EDGPU_MAX_ORB], cd[2 * EDGPU_MAX_ORB * EDGPU_MAX_ORB];
  double v[2 * EDGPU_MAX_ORB * EDGPU_MAX_ORB];
  int m = hd_nonlocal_row(P, (uint32_t)map_up[r % dimup], (uint32_t)map_dw[coloff + r / dimup], cu, cd, v);
  int64_t p = rowptr[r];
  for (int k = 0; k < m; k++) {
    int64_t ju = hd_binary_search(map_up, dimup, (int32_t)cu[k]) - 1;
    int64_t jd = hd_binary_search(map_dw, dimdw, (int32_t)cd[k]) - 1;
    // spin-exchange / pair-hopping targets of one row are pairwise distinct, so no entry of
    // spH0nd ever accumulates (sp_insert_element's append branch only)
    ocols[p + k] = ju + jd * dimup;
    ovals[p + k] = v[k];
  }
}

// Hs%map of one spin species on the device (build_sector, ED_SETUP.f90:764-777) without the factors
int build_sector_map_device(edgpu_ctx *c, int npart, int32_t **d_map) {
  const int64_t n = binom64(c, c->ns, npart);
  CK(cudaMalloc(d_map, (size_t)std::max<int64_t>(n, 1) * sizeof(int32_t)));
  uint64_t top = 1ull << c->ns;
  int blocks = (int)((top + 255) / 256);
  if (blocks > 65535 * 4) blocks = 65535 * 4;
  k_build_map<<<blocks, 256, 0, c->stream>>>(c->ns, npart, c->d_binom, *d_map);
  CKL(c);
  return EDGPU_OK;
}

static int build_factor(edgpu_ctx *c, Factor &f, int spin, int npart, bool stored) {
  f.n = binom64(c, c->ns, npart);
  CK(cudaMalloc(&f.d_map, (size_t)f.n * sizeof(int32_t)));
  {
    uint64_t top = 1ull << c->ns;
    int blocks = (int)((top + 255) / 256);
    if (blocks > 65535 * 4) blocks = 65535 * 4;
    k_build_map<<<blocks, 256, 0, c->stream>>>(c->ns, npart, c->d_binom, f.d_map);
    CKL(c);
  }
  int blocks = (int)((f.n + 127) / 128);
  // the CSR factors are built in both modes: in "direct" mode they stand in for the per-call
  // c/cdg/binary_search regeneration (they are O(DimUp*Ns/2), not O(dim))
  int32_t *d_counts = nullptr;
  CK(cudaMalloc(&d_counts, (size_t)(f.n + 1) * sizeof(int32_t)));
  CK(cudaMemsetAsync(d_counts, 0, (size_t)(f.n + 1) * sizeof(int32_t), c->stream));
  k_factor_count<<<blocks, 128, 0, c->stream>>>(c->dp, spin, f.d_map, f.n, d_counts);
  CKL(c);
  CK(cudaMalloc(&f.d_rowptr, (size_t)(f.n + 1) * sizeof(int32_t)));
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_counts, f.d_rowptr, (int)(f.n + 1), c->stream);
  void *d_tmp = nullptr;
  CK(cudaMalloc(&d_tmp, tmp_bytes));
  cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_counts, f.d_rowptr, (int)(f.n + 1), c->stream);
  c->launches++;
  int32_t nnz = 0;
  CK(cudaMemcpyAsync(&nnz, f.d_rowptr + f.n, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  f.nnz = nnz;
  // max row length (for kernels that want an ELL bound)
  {
    std::vector<int32_t> hc((size_t)f.n);
    CK(cudaMemcpy(hc.data(), d_counts, (size_t)f.n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    f.maxrow = 0;
    for (int64_t i = 0; i < f.n; i++) if (hc[i] > f.maxrow) f.maxrow = hc[i];
  }
  CK(cudaMalloc(&f.d_cols, (size_t)(nnz > 0 ? nnz : 1) * sizeof(int32_t)));
  CK(cudaMalloc(&f.d_vals, (size_t)(nnz > 0 ? nnz : 1) * sizeof(double)));
  k_factor_fill<<<blocks, 128, 0, c->stream>>>(c->dp, spin, f.d_map, f.n, f.d_rowptr, f.d_cols, f.d_vals);
  CKL(c);
  if (!stored) {
    CK(cudaMalloc(&f.d_dfac, (size_t)(f.n + 2) * sizeof(double)));   // +2: 16-byte bulk copies may read one past the end
    CK(cudaMemsetAsync(f.d_dfac, 0, (size_t)(f.n + 2) * sizeof(double), c->stream));
    k_dfac<<<blocks, 128, 0, c->stream>>>(c->dp, spin, f.d_map, f.n, f.d_dfac);
    CKL(c);
  }
  CK(cudaStreamSynchronize(c->stream));
  cudaFree(d_counts);
  cudaFree(d_tmp);
  return EDGPU_OK;
}

static void free_factor(Factor &f) {
  cudaFree(f.d_map); cudaFree(f.d_rowptr); cudaFree(f.d_cols); cudaFree(f.d_vals); cudaFree(f.d_dfac);
  f = Factor();
}

extern "C" int edgpu_build_hv_sector(edgpu_ctx *c, int isector) {
  if (!c) return edgpu_set_err(EDGPU_ERR_INVALID, "ctx == NULL");
  if (c->hstatus) return edgpu_set_err(EDGPU_ERR_INVALID, "build_Hv_sector: a sector is already live (Hstatus=T)");
  CK(cudaSetDevice(c->device));
  if (!c->hp.ed_total_ud) {                                   // ed_buildh_orbs: its own builder and operator (orbs.cu)
    c->isector = isector; c->nup = -1; c->ndw = -1;
    c->hstatus = true;
    int rc = orbs_build(c, isector);
    if (rc) { edgpu_delete_hv_sector(c); return rc; }
    g_current = c;
    return EDGPU_OK;
  }
  int nup, ndw;
  TRY(edgpu_get_nup_ndw(c, isector, &nup, &ndw));
  c->isector = isector; c->nup = nup; c->ndw = ndw;
  c->dimup = binom64(c, c->ns, nup);
  c->dimdw = binom64(c, c->ns, ndw);
  if (c->dimdw < c->nranks)                                 // the reference shrinks the communicator here
    return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "DimDw=%lld < nranks=%d (communicator shrinking, ED_HAMILTONIAN.f90:66-94, is out of scope)",
                         (long long)c->dimdw, c->nranks);
  edgpu_split(c->dimdw, c->nranks, c->rank, &c->qdw, &c->coloff);
  edgpu_split(c->dimup, c->nranks, c->rank, &c->qup, &c->rowoff);
  c->nel = c->dimup * c->qdw;
  c->nloc = c->nel * c->dimph;
  const bool stored = c->hp.ed_sparse_h != 0;
  c->hstatus = true;
  int rc = build_factor(c, c->up, 0, nup, stored);
  if (!rc) rc = build_factor(c, c->dw, 1, ndw, stored);
  if (rc) { edgpu_delete_hv_sector(c); return rc; }
  if (stored) {
    CK(cudaMalloc(&c->d_diag, (size_t)c->nel * sizeof(double)));
    dim3 grid((unsigned)((c->dimup + 255) / 256), (unsigned)(c->qdw < 32768 ? c->qdw : 32768));
    k_diag_stored<<<grid, 256, 0, c->stream>>>(c->dp, c->up.d_map, c->dw.d_map, c->dimup, c->coloff, c->qdw, c->d_diag);
    CKL(c);
  }
  if (c->dp.jhflag) {
    int64_t *d_counts = nullptr;
    CK(cudaMalloc(&d_counts, (size_t)(c->nel + 1) * sizeof(int64_t)));
    CK(cudaMemsetAsync(d_counts, 0, (size_t)(c->nel + 1) * sizeof(int64_t), c->stream));
    if (c->nel + 1 > (int64_t)0x7fffffff * 128)
      return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "spH0nd: local dimension %lld exceeds the builder's grid", (long long)c->nel);
    const unsigned blocks = (unsigned)((c->nel + 127) / 128);
    k_nd_count<<<blocks, 128, 0, c->stream>>>(c->dp, c->up.d_map, c->dw.d_map, c->dimup, c->coloff, c->nel, d_counts);
    CKL(c);
    CK(cudaMalloc(&c->d_nd_rowptr, (size_t)(c->nel + 1) * sizeof(int64_t)));
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_counts, c->d_nd_rowptr, (int64_t)(c->nel + 1), c->stream);
    void *d_tmp = nullptr;
    CK(cudaMalloc(&d_tmp, tmp_bytes));
    cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_counts, c->d_nd_rowptr, (int64_t)(c->nel + 1), c->stream);
    c->launches++;
    CK(cudaMemcpyAsync(&c->nd_nnz, c->d_nd_rowptr + c->nel, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMalloc(&c->d_nd_cols, (size_t)(c->nd_nnz > 0 ? c->nd_nnz : 1) * sizeof(int64_t)));
    CK(cudaMalloc(&c->d_nd_vals, (size_t)(c->nd_nnz > 0 ? c->nd_nnz : 1) * sizeof(double)));
    k_nd_fill<<<blocks, 128, 0, c->stream>>>(c->dp, c->up.d_map, c->dw.d_map, c->dimup, c->dimdw, c->coloff, c->nel,
                                             c->d_nd_rowptr, c->d_nd_cols, c->d_nd_vals);
    CKL(c);
    CK(cudaStreamSynchronize(c->stream));
    cudaFree(d_counts);
    cudaFree(d_tmp);
  }
  CK(cudaStreamSynchronize(c->stream));
  if (c->nranks > 1 && !c->opt_no_peer && !c->dp.jhflag && (c->algo == EDGPU_ALGO_AUTO || c->algo == EDGPU_ALGO_FAST)) {
    // sharded fast path: the halo slab is allocated and mapped here, collectively (the plan is rank independent,
    // so either every rank gets a slab of the same size or none does)
    size_t hb = 0;
    int rc2 = fast_halo_bytes(c, &hb);
    if (!rc2) rc2 = comm_symm_setup(c, hb);
    if (rc2) { edgpu_delete_hv_sector(c); return rc2; }
  }
  g_current = c;
  return EDGPU_OK;
}

extern "C" int edgpu_delete_hv_sector(edgpu_ctx *c) {
  if (!c) return edgpu_set_err(EDGPU_ERR_INVALID, "ctx == NULL");
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  tiled_plan_free(c);
  fast_plan_free(c);
  orbs_free(c);
  free_factor(c->up);
  free_factor(c->dw);
  cudaFree(c->d_diag); c->d_diag = nullptr;
  cudaFree(c->d_nd_rowptr); cudaFree(c->d_nd_cols); cudaFree(c->d_nd_vals);
  c->d_nd_rowptr = c->d_nd_cols = nullptr; c->d_nd_vals = nullptr; c->nd_nnz = 0;
  double **bufs[] = { &c->d_in, &c->d_out, &c->d_vt, &c->d_hvt, &c->d_send, &c->d_recv, &c->d_full,
                      &c->d_lx, &c->d_lp, &c->d_lt, &c->d_l0, &c->d_lv };
  for (auto b : bufs) vec_free(c, b);
  comm_symm_teardown(c);
  c->hstatus = false;
  c->isector = 0;
  if (g_current == c) g_current = nullptr;
  return EDGPU_OK;
}

extern "C" int edgpu_destroy(edgpu_ctx *c) {
  if (!c) return EDGPU_OK;
  cudaSetDevice(c->device);
  if (c->hstatus) edgpu_delete_hv_sector(c);
  edgpu_comm_finalize(c);
  cudaFree(c->d_gs);
  cudaFree(c->d_binom); cudaFree(c->d_partials); cudaFree(c->d_st);
  cudaFree(c->d_alanc); cudaFree(c->d_blanc);
  cudaFreeHost(c->h_pinned);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->stream2) cudaStreamDestroy(c->stream2);
  for (int i = 0; i < 8; i++) if (c->pev[i]) cudaEventDestroy(c->pev[i]);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return EDGPU_OK;
}

// ------------------------------------------------------------------------------------------
// the operator through host pointers (spHtimesV_p)
// ------------------------------------------------------------------------------------------
extern "C" int edgpu_hxv_device(edgpu_ctx *c, int64_t nloc, const double *d_v, double *d_hv) {
  if (!c || !c->hstatus) return edgpu_set_err(EDGPU_ERR_INVALID, "HxV: Hsector NOT set");
  if (nloc != c->nloc) return edgpu_set_err(EDGPU_ERR_INVALID, "HxV: Nloc=%lld != vecDim=%lld", (long long)nloc, (long long)c->nloc);
  CK(cudaSetDevice(c->device));
  return hxv_apply(c, d_v, d_hv);                               // nobody but this rank ever reads d_v (halo is pushed)
}

extern "C" int edgpu_hxv(edgpu_ctx *c, int64_t nloc, const double *v, double *hv) {
  if (!c || !c->hstatus) return edgpu_set_err(EDGPU_ERR_INVALID, "HxV: Hsector NOT set");
  if (nloc != c->nloc) return edgpu_set_err(EDGPU_ERR_INVALID, "HxV: Nloc=%lld != vecDim=%lld", (long long)nloc, (long long)c->nloc);
  CK(cudaSetDevice(c->device));
  TRY(vec_alloc(c, &c->d_in, c->nloc));
  TRY(vec_alloc(c, &c->d_out, c->nloc));
  CK(cudaMemcpyAsync(c->d_in, v, (size_t)nloc * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  TRY(hxv_apply(c, c->d_in, c->d_out));
  CK(cudaMemcpyAsync(hv, c->d_out, (size_t)nloc * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return EDGPU_OK;
}

extern "C" void edgpu_sphtimesv(const int32_t *nloc, const double *v, double *hv) {
  if (!g_current) { fprintf(stderr, "edgpu_sphtimesv ERROR: Hsector NOT set\n"); abort(); }
  int rc = edgpu_hxv(g_current, (int64_t)*nloc, v, hv);
  if (rc) { fprintf(stderr, "edgpu_sphtimesv ERROR: %s\n", edgpu_last_error()); abort(); }
}

// ------------------------------------------------------------------------------------------
// introspection
// ------------------------------------------------------------------------------------------
extern "C" int edgpu_get_dims(const edgpu_ctx *c, int64_t *dimup, int64_t *dimdw, int64_t *qdw,
                              int64_t *ishift, int64_t *nloc) {
  if (!c || !c->hstatus) return edgpu_set_err(EDGPU_ERR_INVALID, "no live sector");
  if (dimup) *dimup = c->dimup;
  if (dimdw) *dimdw = c->dimdw;
  if (qdw) *qdw = c->qdw;
  if (ishift) *ishift = c->coloff * c->dimup;               // mpiIshift, ED_HAMILTONIAN.f90:110
  if (nloc) *nloc = c->nloc;
  return EDGPU_OK;
}
extern "C" int edgpu_get_sector_map(const edgpu_ctx *c, int which, int32_t *out) {
  if (!c || !c->hstatus) return edgpu_set_err(EDGPU_ERR_INVALID, "no live sector");
  if (c->orbs) return edgpu_set_err(EDGPU_ERR_INVALID, "ed_total_ud = F: use edgpu_get_orbs_factor");
  const Factor &f = which ? c->dw : c->up;
  CK(cudaMemcpy(out, f.d_map, (size_t)f.n * sizeof(int32_t), cudaMemcpyDeviceToHost));
  return EDGPU_OK;
}
extern "C" int edgpu_get_csr(const edgpu_ctx *c, int which, int64_t *nrow, int64_t *nnz,
                             int64_t *rowptr, int64_t *cols, double *vals) {
  if (!c || !c->hstatus) return edgpu_set_err(EDGPU_ERR_INVALID, "no live sector");
  if (c->orbs) return edgpu_set_err(EDGPU_ERR_INVALID, "ed_total_ud = F: use edgpu_get_orbs_factor");
  if (which == 2) {
    if (nrow) *nrow = c->dp.jhflag ? c->nel : 0;
    if (nnz) *nnz = c->nd_nnz;
    if (!rowptr || !c->dp.jhflag) return EDGPU_OK;
    CK(cudaMemcpy(rowptr, c->d_nd_rowptr, (size_t)(c->nel + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cols, c->d_nd_cols, (size_t)c->nd_nnz * sizeof(int64_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(vals, c->d_nd_vals, (size_t)c->nd_nnz * sizeof(double), cudaMemcpyDeviceToHost));
    return EDGPU_OK;
  }
  const Factor &f = which ? c->dw : c->up;
  if (nrow) *nrow = f.n;
  if (nnz) *nnz = f.nnz;
  if (!rowptr) return EDGPU_OK;
  std::vector<int32_t> rp((size_t)f.n + 1), cc((size_t)f.nnz);
  CK(cudaMemcpy(rp.data(), f.d_rowptr, rp.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(cc.data(), f.d_cols, cc.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < rp.size(); i++) rowptr[i] = rp[i];
  for (size_t i = 0; i < cc.size(); i++) cols[i] = cc[i];
  CK(cudaMemcpy(vals, f.d_vals, (size_t)f.nnz * sizeof(double), cudaMemcpyDeviceToHost));
  return EDGPU_OK;
}

__global__ void k_diag_direct(DevParams P, const int32_t *__restrict__ map_up, const int32_t *__restrict__ map_dw,
                              const double *__restrict__ fu, const double *__restrict__ fd,
                              int64_t dimup, int64_t coloff, int64_t qdw, double *__restrict__ out) {
  for (int64_t jl = blockIdx.y; jl < qdw; jl += gridDim.y) {
    uint32_t mdw = (uint32_t)map_dw[coloff + jl];
    double dj = fd[coloff + jl];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < dimup; i += (int64_t)gridDim.x * blockDim.x)
      out[i + jl * dimup] = fu[i] + dj + hd_diag_cross(P, (uint32_t)map_up[i], mdw);
  }
}
extern "C" int edgpu_get_diag(const edgpu_ctx *cc, double *out, int64_t nloc) {
  edgpu_ctx *c = const_cast<edgpu_ctx *>(cc);
  if (!c || !c->hstatus) return edgpu_set_err(EDGPU_ERR_INVALID, "no live sector");
  if (nloc != c->nel) return edgpu_set_err(EDGPU_ERR_INVALID, "get_diag: nloc is the electron part DimUp*mpiQdw");
  if (c->d_diag) {
    CK(cudaMemcpy(out, c->d_diag, (size_t)nloc * sizeof(double), cudaMemcpyDeviceToHost));
    return EDGPU_OK;
  }
  double *d_tmp = nullptr;
  CK(cudaMalloc(&d_tmp, (size_t)nloc * sizeof(double)));
  dim3 grid((unsigned)((c->dimup + 255) / 256), (unsigned)(c->qdw < 32768 ? c->qdw : 32768));
  k_diag_direct<<<grid, 256, 0, c->stream>>>(c->dp, c->up.d_map, c->dw.d_map, c->up.d_dfac, c->dw.d_dfac,
                                             c->dimup, c->coloff, c->qdw, d_tmp);
  CKL(c);
  CK(cudaMemcpyAsync(out, d_tmp, (size_t)nloc * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  cudaFree(d_tmp);
  return EDGPU_OK;
}

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
extern "C" int edgpu_dev_alloc(edgpu_ctx *c, int64_t nbytes, void **dptr) {
  if (!c || !dptr) return edgpu_set_err(EDGPU_ERR_INVALID, "ctx/dptr == NULL");
  CK(cudaSetDevice(c->device));
  CK(cudaMalloc(dptr, (size_t)nbytes + 16));
  return EDGPU_OK;
}
extern "C" int edgpu_dev_free(edgpu_ctx *c, void *dptr) {
  if (!c) return edgpu_set_err(EDGPU_ERR_INVALID, "ctx == NULL");
  if (!dptr) return EDGPU_OK;
  CK(cudaSetDevice(c->device));
  CK(cudaFree(dptr));
  return EDGPU_OK;
}
extern "C" int edgpu_dev_upload(edgpu_ctx *c, void *dptr, const void *host, int64_t nbytes) {
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpyAsync(dptr, host, (size_t)nbytes, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return EDGPU_OK;
}
extern "C" int edgpu_dev_download(edgpu_ctx *c, void *host, const void *dptr, int64_t nbytes) {
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpyAsync(host, dptr, (size_t)nbytes, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return EDGPU_OK;
}
__global__ void k_fill_bench(double *__restrict__ v, int64_t n, int64_t off) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    v[i] = sin(0.37 * (double)(off + i + 1)) + 0.1;         // v_i = sin(0.37 i) + 0.1
}
extern "C" int edgpu_dev_fill_bench_vector(edgpu_ctx *c, double *d_v, int64_t nloc, int64_t global_offset) {
  CK(cudaSetDevice(c->device));
  k_fill_bench<<<c->sm_count * 8, 256, 0, c->stream>>>(d_v, nloc, global_offset);
  CKL(c);
  return EDGPU_OK;
}
extern "C" int edgpu_sync(edgpu_ctx *c) {
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  return EDGPU_OK;
}
extern "C" int edgpu_launch_count(const edgpu_ctx *c, int64_t *n) {
  *n = c->launches;
  return EDGPU_OK;
}
// Per-kernel split of a device-resident H*v (single rank): CUDA events between the passes, summed over
// `reps` applications.  ms_pass[6]; names receives up to 6 NUL-terminated kernel names of 32 bytes each.
extern "C" int edgpu_time_hxv_passes(edgpu_ctx *c, int64_t nloc, const double *d_v, double *d_hv, int reps,
                                     int *npasses, double *ms_pass, char *names) {
  if (!c || !c->hstatus) return edgpu_set_err(EDGPU_ERR_INVALID, "HxV: Hsector NOT set");
  if (nloc != c->nloc) return edgpu_set_err(EDGPU_ERR_INVALID, "nloc mismatch");
  CK(cudaSetDevice(c->device));
  for (int i = 0; i < 8; i++) if (!c->pev[i]) CK(cudaEventCreate(&c->pev[i]));
  for (int i = 0; i < 6; i++) ms_pass[i] = 0.0;
  int np = 0;
  for (int r = 0; r < reps; r++) {
    c->prof = true; c->prof_n = 0;
    int rc = hxv_apply(c, d_v, d_hv);
    if (!rc) prof_mark(c, "end");
    c->prof = false;
    if (rc) return rc;
    CK(cudaStreamSynchronize(c->stream));
    np = c->prof_n - 1;
    for (int i = 0; i < np && i < 6; i++) {
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, c->pev[i], c->pev[i + 1]));
      ms_pass[i] += ms;
      if (names) { strncpy(names + 32 * i, c->prof_name[i], 31); names[32 * i + 31] = 0; }
    }
  }
  if (npasses) *npasses = np < 6 ? np : 6;
  return EDGPU_OK;
}
extern "C" int edgpu_time_hxv_device(edgpu_ctx *c, int64_t nloc, const double *d_v, double *d_hv,
                                     int reps, double *ms_total) {
  if (!c || !c->hstatus) return edgpu_set_err(EDGPU_ERR_INVALID, "HxV: Hsector NOT set");
  if (nloc != c->nloc) return edgpu_set_err(EDGPU_ERR_INVALID, "nloc mismatch");
  CK(cudaSetDevice(c->device));
  CK(cudaEventRecord(c->ev0, c->stream));
  for (int r = 0; r < reps; r++) TRY(hxv_apply(c, d_v, d_hv));
  CK(cudaEventRecord(c->ev1, c->stream));
  CK(cudaEventSynchronize(c->ev1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  *ms_total = ms;
  return EDGPU_OK;
}
