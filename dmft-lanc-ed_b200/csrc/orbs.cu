// orbs.cu -- the orbital-resolved operator (ed_total_ud = F): one (Nup, Ndw) pair per orbital.
//
// Replaces ed_buildh_orbs / spMatVec_orbs / directMatVec_orbs (ED_HAMILTONIAN_SPARSE_HxV.f90:206-370, 487-564;
// ED_HAMILTONIAN_DIRECT_HxV.f90) with stored/Orbs/H_local.f90, H_up.f90, H_dw.f90.  The Fock space factorises into
// 2*Norb words of Ns_Orb = 1 + Nbath bits (up word of orbital 1..Norb, then the dw words), every word with its own
// conserved particle number; the sector vector is the dense tensor over those 2*Norb indices, first index fastest
// (state2indices, ED_SETUP.f90:520-545).  H = Hd + sum_f (I x ... x H_f x ... x I) with H_f the single-band star of
// that orbital and spin (impurity = bit 0 of the word, bath level k = bit k), so every factor has the structure of
// the single-band case; the diagonal couples all words (Kanamori density-density terms) and is evaluated from the
// occupations reordered to the Ns site numbering (breorder, ED_SETUP.f90:963-979) with the reference's summation order.
//
// One kernel, one pass: a thread owns an element, decodes its 2*Norb indices and gathers the <= Nbath hops of every
// factor along that factor's stride.  Single rank only (the reference splits the combined dw index; not built).
#include <algorithm>
#include <vector>

#include "engine.h"

struct OrbsPlan {
  int nfac = 0, nso = 0;
  int nq[2 * EDGPU_MAX_ORB] = {0};
  int64_t dims[2 * EDGPU_MAX_ORB] = {0}, stride[2 * EDGPU_MAX_ORB] = {0};
  int32_t *d_map[2 * EDGPU_MAX_ORB] = {nullptr};
  int32_t *d_rowptr[2 * EDGPU_MAX_ORB] = {nullptr}, *d_cols[2 * EDGPU_MAX_ORB] = {nullptr};
  double *d_vals[2 * EDGPU_MAX_ORB] = {nullptr};
  std::vector<int32_t> h_map[2 * EDGPU_MAX_ORB], h_rowptr[2 * EDGPU_MAX_ORB], h_cols[2 * EDGPU_MAX_ORB];
  std::vector<double> h_vals[2 * EDGPU_MAX_ORB];
  double *d_diag = nullptr;      // spH0d (stored mode)
};

struct OrbsArgs {
  int nfac, norb, nbath;
  int64_t n;
  int64_t dims[2 * EDGPU_MAX_ORB], stride[2 * EDGPU_MAX_ORB];
  const int32_t *map[2 * EDGPU_MAX_ORB], *rowptr[2 * EDGPU_MAX_ORB], *cols[2 * EDGPU_MAX_ORB];
  const double *vals[2 * EDGPU_MAX_ORB];
  const double *diag;            // nullptr: recompute (direct mode)
  const double *x;
  double *y;
};

// the Ns-bit up and dw words of element i (breorder): impurity orbital io -> bit io, bath level k of orbital io ->
// bit Norb + io*Nbath + k - 1 (getBathStride, normal bath)
__device__ __forceinline__ void orbs_words(const OrbsArgs &a, const int64_t *idx, uint32_t *mup, uint32_t *mdw) {
  uint32_t mu = 0, md = 0;
  for (int f = 0; f < a.nfac; f++) {
    const uint32_t w = (uint32_t)a.map[f][idx[f]];
    const int io = f % a.norb;
    const uint32_t full = ((w & 1u) << io) | ((w >> 1) << (a.norb + io * a.nbath));
    if (f < a.norb) mu |= full; else md |= full;
  }
  *mup = mu; *mdw = md;
}
__global__ void k_orbs_diag(DevParams P, OrbsArgs a, double *__restrict__ diag) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t idx[2 * EDGPU_MAX_ORB], c = i;
    for (int f = 0; f < a.nfac; f++) { idx[f] = c % a.dims[f]; c /= a.dims[f]; }
    uint32_t mu, md;
    orbs_words(a, idx, &mu, &md);
    diag[i] = hd_diag_element(P, mu, md);
  }
}
// y = H x, spMatVec_orbs order: diagonal, then per orbital the up entries followed by the dw entries
__global__ void __launch_bounds__(256) k_hxv_orbs(DevParams P, OrbsArgs a) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t idx[2 * EDGPU_MAX_ORB], c = i;
    for (int f = 0; f < a.nfac; f++) { idx[f] = c % a.dims[f]; c /= a.dims[f]; }
    double d;
    if (a.diag) d = a.diag[i];
    else { uint32_t mu, md; orbs_words(a, idx, &mu, &md); d = hd_diag_element(P, mu, md); }
    double acc = d * a.x[i];
    for (int iud = 0; iud < a.norb; iud++)
      for (int sp = 0; sp < 2; sp++) {
        const int f = iud + sp * a.norb;
        const int32_t r = (int32_t)idx[f];
        for (int32_t p = a.rowptr[f][r]; p < a.rowptr[f][r + 1]; p++)
          acc += a.vals[f][p] * a.x[i + (int64_t)(a.cols[f][p] - r) * a.stride[f]];
      }
    a.y[i] = acc;
  }
}

// ---- sector numbering, ED_SETUP.f90:446-500 with QN = [Nups, Ndws], factor Ns_Orb + 1 -------------------------------
extern "C" int edgpu_get_sector_orbs(const edgpu_ctx *c, const int *nups, const int *ndws, int *isector) {
  if (!c || !nups || !ndws || !isector) return edgpu_set_err(EDGPU_ERR_INVALID, "get_sector_orbs: NULL argument");
  const int norb = c->dp.norb, nso = c->dp.nbath + 1, nind = 2 * norb;
  int64_t s = 1;
  for (int i = nind; i >= 1; i--) {
    const int qn = (i <= norb) ? nups[i - 1] : ndws[i - 1 - norb];
    if (qn < 0 || qn > nso) return edgpu_set_err(EDGPU_ERR_INVALID, "get_sector_orbs: occupation out of range");
    int64_t pw = 1;
    for (int k = 0; k < nind - i; k++) pw *= (nso + 1);
    s += qn * pw;
  }
  if (s > 0x7fffffff) return edgpu_set_err(EDGPU_ERR_INVALID, "get_sector_orbs: sector number overflows");
  *isector = (int)s;
  return EDGPU_OK;
}
extern "C" int edgpu_get_qn_orbs(const edgpu_ctx *c, int isector, int *nups, int *ndws) {
  if (!c || !nups || !ndws) return edgpu_set_err(EDGPU_ERR_INVALID, "get_qn_orbs: NULL argument");
  const int norb = c->dp.norb, nso = c->dp.nbath + 1, nind = 2 * norb;
  int64_t nsec = 1;
  for (int k = 0; k < nind; k++) nsec *= (nso + 1);
  if (isector < 1 || isector > nsec) return edgpu_set_err(EDGPU_ERR_INVALID, "isector out of range");
  int count = isector - 1, ind[2 * EDGPU_MAX_ORB];
  for (int i = 0; i < nind; i++) { ind[i] = count % (nso + 1); count /= (nso + 1); }   // get_Nup / get_Ndw
  for (int k = 0; k < norb; k++) { nups[k] = ind[nind - 1 - k]; ndws[k] = ind[norb - 1 - k]; }
  return EDGPU_OK;
}

int orbs_free(edgpu_ctx *c) {
  OrbsPlan *p = c->orbs;
  if (!p) return EDGPU_OK;
  for (int f = 0; f < 2 * EDGPU_MAX_ORB; f++) { cudaFree(p->d_map[f]); cudaFree(p->d_rowptr[f]); cudaFree(p->d_cols[f]); cudaFree(p->d_vals[f]); }
  cudaFree(p->d_diag);
  delete p;
  c->orbs = nullptr;
  return EDGPU_OK;
}

template <typename T>
static int up(T **d, const std::vector<T> &h) {
  CK(cudaMalloc(d, std::max<size_t>(h.size(), 1) * sizeof(T)));
  if (!h.empty()) CK(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return EDGPU_OK;
}

static void fill_args(const edgpu_ctx *c, OrbsArgs &a) {
  const OrbsPlan *p = c->orbs;
  a.nfac = p->nfac; a.norb = c->dp.norb; a.nbath = c->dp.nbath; a.n = c->nloc;
  for (int f = 0; f < p->nfac; f++) {
    a.dims[f] = p->dims[f]; a.stride[f] = p->stride[f];
    a.map[f] = p->d_map[f]; a.rowptr[f] = p->d_rowptr[f]; a.cols[f] = p->d_cols[f]; a.vals[f] = p->d_vals[f];
  }
  a.diag = p->d_diag;
}

// One factor on the host (pure arithmetic, also reachable without a GPU through edgpu_selftest_orbs_factor): Hs(f)%map
// = the ascending words of Ns_Orb bits with n set (build_sector, ED_SETUP.f90:764-777) and spH0ups(iorb) / spH0dws(iorb)
// in the reference's insertion order -- outer loop over the source state, bath level inner (stored/Orbs/H_up.f90).
void orbs_factor_host(const DevParams &dp, int f, int n, std::vector<int32_t> &map, std::vector<int32_t> &rowptr,
                      std::vector<int32_t> &cols, std::vector<double> &vals) {
  const int norb = dp.norb, nbath = dp.nbath, nso = nbath + 1;
  map.clear(); rowptr.clear(); cols.clear(); vals.clear();
  for (uint32_t w = 0; w < (1u << nso); w++) if (__builtin_popcount(w) == n) map.push_back((int32_t)w);
  const int64_t dim = (int64_t)map.size();
  const int io = f % norb;
  const double *v = (f < norb) ? dp.bv_up : dp.bv_dw;
  std::vector<std::vector<int32_t>> rc((size_t)dim);
  std::vector<std::vector<double>> rv((size_t)dim);
  for (int64_t j = 0; j < dim; j++) {
    const uint32_t m = (uint32_t)map[(size_t)j];
    for (int kp = 1; kp <= nbath; kp++) {
      const double vk = v[io * nbath + kp - 1];
      if (vk == 0.0) continue;
      const bool imp = (m & 1u) != 0, bath = ((m >> kp) & 1u) != 0;
      if (imp == bath) continue;
      const uint32_t k2 = m ^ 1u ^ (1u << kp);
      // imp -> bath: c(1) has no sites below it, cdg(1+kp) counts the occupied sites below it on the state without
      // the impurity electron; bath -> imp: c(1+kp) counts them on m (impurity empty), cdg(1) none
      const uint32_t below = (imp ? (m & ~1u) : m) & ((1u << kp) - 1u);
      const double sg = (__builtin_popcount(below) & 1) ? -1.0 : 1.0;
      const int64_t i = std::lower_bound(map.begin(), map.end(), (int32_t)k2) - map.begin();   // binary_search
      rc[(size_t)i].push_back((int32_t)j);
      rv[(size_t)i].push_back(vk * sg);
    }
  }
  rowptr.assign(1, 0);
  for (int64_t i = 0; i < dim; i++) {
    cols.insert(cols.end(), rc[(size_t)i].begin(), rc[(size_t)i].end());
    vals.insert(vals.end(), rv[(size_t)i].begin(), rv[(size_t)i].end());
    rowptr.push_back((int32_t)cols.size());
  }
}

// build_Hv_sector for ed_total_ud = F.  The factors are tiny (C(Ns_Orb, n) rows): built on the host in the reference's
// insertion order (outer loop over the source state, bath level inner; stored/Orbs/H_up.f90), then uploaded.
int orbs_build(edgpu_ctx *c, int isector) {
  if (c->nranks != 1) return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "ed_total_ud = F on more than one rank is not built");
  const int norb = c->dp.norb, nbath = c->dp.nbath, nso = nbath + 1;
  int nups[EDGPU_MAX_ORB], ndws[EDGPU_MAX_ORB];
  TRY(edgpu_get_qn_orbs(c, isector, nups, ndws));
  OrbsPlan *p = new OrbsPlan();
  c->orbs = p;
  p->nfac = 2 * norb; p->nso = nso;
  int64_t dim = 1;
  for (int f = 0; f < p->nfac; f++) {
    const int n = (f < norb) ? nups[f] : ndws[f - norb];
    p->nq[f] = n;
    p->dims[f] = (int64_t)c->h_binom[nso * EDGPU_BINOM_LD + n];
    p->stride[f] = dim;
    dim *= p->dims[f];
    if (dim > ((int64_t)1 << 40)) return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "sector too large");
    orbs_factor_host(c->dp, f, n, p->h_map[f], p->h_rowptr[f], p->h_cols[f], p->h_vals[f]);
    TRY(up(&p->d_map[f], p->h_map[f]));
    TRY(up(&p->d_rowptr[f], p->h_rowptr[f]));
    TRY(up(&p->d_cols[f], p->h_cols[f]));
    TRY(up(&p->d_vals[f], p->h_vals[f]));
  }
  c->dimup = 1; c->dimdw = 1;
  for (int f = 0; f < norb; f++) { c->dimup *= p->dims[f]; c->dimdw *= p->dims[f + norb]; }
  c->qdw = c->dimdw; c->coloff = 0; c->qup = c->dimup; c->rowoff = 0;
  c->nloc = dim; c->nel = dim;
  if (c->hp.ed_sparse_h) {
    CK(cudaMalloc(&p->d_diag, (size_t)std::max<int64_t>(dim, 1) * sizeof(double)));
    OrbsArgs a{};
    fill_args(c, a);
    a.diag = nullptr;
    const int grid = (int)std::min<int64_t>((dim + 255) / 256, (int64_t)c->sm_count * 16);
    k_orbs_diag<<<std::max(grid, 1), 256, 0, c->stream>>>(c->dp, a, p->d_diag);
    CKL(c);
  }
  CK(cudaStreamSynchronize(c->stream));
  return EDGPU_OK;
}

int orbs_apply(edgpu_ctx *c, const double *d_x, double *d_y) {
  OrbsArgs a{};
  fill_args(c, a);
  a.x = d_x; a.y = d_y;
  const int grid = (int)std::min<int64_t>((c->nloc + 255) / 256, (int64_t)c->sm_count * 16);
  prof_mark(c, "k_hxv_orbs");
  k_hxv_orbs<<<std::max(grid, 1), 256, 0, c->stream>>>(c->dp, a);
  CKL(c);
  return EDGPU_OK;
}

// introspection (bit-exact checks): factor f < Norb = up word of orbital f+1, else dw word of orbital f+1-Norb.
// dims[2*Norb]; any array may be NULL; rowptr/cols are 0-based.
extern "C" int edgpu_get_orbs_dims(const edgpu_ctx *c, int64_t *dims, int64_t *dim) {
  if (!c || !c->hstatus || !c->orbs) return edgpu_set_err(EDGPU_ERR_INVALID, "no live ed_total_ud = F sector");
  for (int f = 0; f < c->orbs->nfac; f++) dims[f] = c->orbs->dims[f];
  if (dim) *dim = c->nloc;
  return EDGPU_OK;
}
extern "C" int edgpu_get_orbs_factor(const edgpu_ctx *c, int f, int32_t *map, int64_t *nnz, int64_t *rowptr, int64_t *cols, double *vals) {
  if (!c || !c->hstatus || !c->orbs) return edgpu_set_err(EDGPU_ERR_INVALID, "no live ed_total_ud = F sector");
  const OrbsPlan *p = c->orbs;
  if (f < 0 || f >= p->nfac) return edgpu_set_err(EDGPU_ERR_INVALID, "factor index out of range");
  if (map) for (size_t k = 0; k < p->h_map[f].size(); k++) map[k] = p->h_map[f][k];
  if (nnz) *nnz = (int64_t)p->h_cols[f].size();
  if (rowptr) for (size_t k = 0; k < p->h_rowptr[f].size(); k++) rowptr[k] = p->h_rowptr[f][k];
  if (cols) for (size_t k = 0; k < p->h_cols[f].size(); k++) cols[k] = p->h_cols[f][k];
  if (vals) for (size_t k = 0; k < p->h_vals[f].size(); k++) vals[k] = p->h_vals[f][k];
  return EDGPU_OK;
}
extern "C" int edgpu_get_orbs_diag(const edgpu_ctx *c, double *out) {
  if (!c || !c->hstatus || !c->orbs || !c->orbs->d_diag) return edgpu_set_err(EDGPU_ERR_INVALID, "no stored diagonal of an ed_total_ud = F sector");
  CK(cudaMemcpy(out, c->orbs->d_diag, (size_t)c->nloc * sizeof(double), cudaMemcpyDeviceToHost));
  return EDGPU_OK;
}
