// hxv.cu -- the sector operator y = H x on the local shard.
//
// Replaces spMatVec_main / spMatVec_MPI_main (ED_HAMILTONIAN_SPARSE_HxV.f90:391-485, 568-694)
// and directMatVec_main / directMatVec_MPI_main (ED_HAMILTONIAN_DIRECT_HxV.f90:21-95, 180-284).
// The vector is the column-major matrix x(i_up, i_dw):
//     y = Hd o x  +  Hup x  +  x Hdw^T  (+ Hnd x)
// This file holds the one-pass GATHER kernels (every output element gathers its 1+Ns inputs
// straight from global memory / L2); hxv_tiled.cu holds the shared-memory staged kernels that
// the AUTO policy prefers for large single-band sectors.  Both produce the same numbers up to
// summation order.
#include "engine.h"

struct GatherArgs {
  int64_t nfac;                 // length of the contiguous (factor) dimension
  int64_t ncols;                // local columns
  int64_t coloff;               // global index of the first local column
  const double *x;
  double *y;
  // diagonal
  const double *diag;           // stored: one value per local element
  const double *dfac_c, *dfac_s;// direct: factor tables (contiguous dim, strided dim)
  const int32_t *map_c, *map_s;
  // contiguous-dimension factor (acts inside a column)
  const int32_t *c_rowptr, *c_cols;
  const double *c_vals;
  // strided-dimension factor (acts across columns, same row); only valid when all columns are local
  const int32_t *s_rowptr, *s_cols;
  const double *s_vals;
  // non-local terms, global column indices into xfull
  const int64_t *nd_rowptr, *nd_cols;
  const double *nd_vals;
  const double *xfull;
  int norb;
  double uloc[EDGPU_MAX_ORB];
  double ust;
};

// DIAG: 0 none, 1 stored, 2 recomputed from the factor tables and the impurity bits
template <int DIAG, bool CFAC, bool SFAC, bool ND, bool ACC>
__global__ void __launch_bounds__(256) k_hxv_gather(GatherArgs a) {
  for (int64_t jl = blockIdx.y; jl < a.ncols; jl += gridDim.y) {
    const double *xc = a.x + jl * a.nfac;
    double *yc = a.y + jl * a.nfac;
    const int64_t jg = a.coloff + jl;
    double dcol = 0.0;
    uint32_t ms = 0;
    int s0 = 0, s1 = 0;
    if (DIAG == 2) { dcol = a.dfac_s[jg]; ms = (uint32_t)a.map_s[jg]; }
    if (SFAC) { s0 = a.s_rowptr[jg]; s1 = a.s_rowptr[jg + 1]; }
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.nfac; i += (int64_t)gridDim.x * blockDim.x) {
      double acc = 0.0;
      if (DIAG == 1) acc = a.diag[i + jl * a.nfac] * xc[i];
      if (DIAG == 2) {
        uint32_t mc = (uint32_t)a.map_c[i];
        double d = a.dfac_c[i] + dcol;
        for (int o = 0; o < a.norb; o++)
          if ((mc >> o) & 1u)
            for (int b = 0; b < a.norb; b++)
              if ((ms >> b) & 1u) d += (o == b) ? a.uloc[o] : a.ust;
        acc = d * xc[i];
      }
      if (CFAC) {
        const int p1 = a.c_rowptr[i + 1];
        for (int p = a.c_rowptr[i]; p < p1; p++) acc += a.c_vals[p] * xc[a.c_cols[p]];
      }
      if (SFAC) {
        for (int p = s0; p < s1; p++) acc += a.s_vals[p] * a.x[i + (int64_t)a.s_cols[p] * a.nfac];
      }
      if (ND) {
        const int64_t r = i + jl * a.nfac;
        const int64_t p1 = a.nd_rowptr[r + 1];
        for (int64_t p = a.nd_rowptr[r]; p < p1; p++) acc += a.nd_vals[p] * a.xfull[a.nd_cols[p]];
      }
      if (ACC) yc[i] += acc; else yc[i] = acc;
    }
  }
}

static void fill_common(const edgpu_ctx *c, GatherArgs &a) {
  a.norb = c->dp.norb;
  for (int i = 0; i < EDGPU_MAX_ORB; i++) a.uloc[i] = c->dp.uloc[i];
  a.ust = c->dp.ust;
}

static dim3 gather_grid(const edgpu_ctx *c, int64_t nfac, int64_t ncols) {
  unsigned gx = (unsigned)((nfac + 255) / 256);
  if (gx > 64) gx = 64;
  int64_t want = (int64_t)c->sm_count * 16 / gx + 1;
  unsigned gy = (unsigned)(ncols < want ? ncols : want);
  if (gy < 1) gy = 1;
  if (gy > 65535) gy = 65535;
  return dim3(gx, gy);
}

template <int DIAG, bool CFAC, bool SFAC>
static void launch_gather(edgpu_ctx *c, const GatherArgs &a, bool nd, bool acc) {
  dim3 g = gather_grid(c, a.nfac, a.ncols);
  if (nd) {
    if (acc) k_hxv_gather<DIAG, CFAC, SFAC, true, true><<<g, 256, 0, c->stream>>>(a);
    else k_hxv_gather<DIAG, CFAC, SFAC, true, false><<<g, 256, 0, c->stream>>>(a);
  } else {
    if (acc) k_hxv_gather<DIAG, CFAC, SFAC, false, true><<<g, 256, 0, c->stream>>>(a);
    else k_hxv_gather<DIAG, CFAC, SFAC, false, false><<<g, 256, 0, c->stream>>>(a);
  }
}

// local part on the (i_up, local i_dw) shard: diagonal + H_up (+ H_dw when every column is local)
static int gather_local(edgpu_ctx *c, const double *d_x, double *d_y, bool with_dw, const double *d_full) {
  GatherArgs a{};
  fill_common(c, a);
  a.nfac = c->dimup; a.ncols = c->qdw; a.coloff = c->coloff;
  a.x = d_x; a.y = d_y;
  a.diag = c->d_diag;
  a.dfac_c = c->up.d_dfac; a.dfac_s = c->dw.d_dfac;
  a.map_c = c->up.d_map; a.map_s = c->dw.d_map;
  a.c_rowptr = c->up.d_rowptr; a.c_cols = c->up.d_cols; a.c_vals = c->up.d_vals;
  a.s_rowptr = c->dw.d_rowptr; a.s_cols = c->dw.d_cols; a.s_vals = c->dw.d_vals;
  a.nd_rowptr = c->d_nd_rowptr; a.nd_cols = c->d_nd_cols; a.nd_vals = c->d_nd_vals;
  a.xfull = d_full;
  const bool nd = c->dp.jhflag && d_full != nullptr;
  const bool stored = c->d_diag != nullptr;
  if (stored) {
    if (with_dw) launch_gather<1, true, true>(c, a, nd, false);
    else launch_gather<1, true, false>(c, a, nd, false);
  } else {
    if (with_dw) launch_gather<2, true, true>(c, a, nd, false);
    else launch_gather<2, true, false>(c, a, nd, false);
  }
  CKL(c);
  return EDGPU_OK;
}

// H_dw on the transposed shard vt(i_dw, local i_up): a contiguous-dimension factor, no diagonal
static int gather_transposed_dw(edgpu_ctx *c, const double *d_vt, double *d_hvt) {
  GatherArgs a{};
  fill_common(c, a);
  a.nfac = c->dimdw; a.ncols = c->qup; a.coloff = c->rowoff;
  a.x = d_vt; a.y = d_hvt;
  a.c_rowptr = c->dw.d_rowptr; a.c_cols = c->dw.d_cols; a.c_vals = c->dw.d_vals;
  launch_gather<0, true, false>(c, a, false, false);
  CKL(c);
  return EDGPU_OK;
}

static int ensure(double **p, int64_t n) {
  if (*p) return EDGPU_OK;
  CK(cudaMalloc(p, (size_t)(n > 0 ? n : 1) * sizeof(double)));
  return EDGPU_OK;
}

// algorithm actually used for the local full operator / for one whole-column factor application
static int pick_local(edgpu_ctx *c) {
  int algo = c->algo;
  if (algo == EDGPU_ALGO_AUTO) algo = fast_supported_local(c) ? EDGPU_ALGO_FAST : (tiled_supported(c) ? EDGPU_ALGO_TILED : EDGPU_ALGO_GATHER);
  return algo;
}
static int pick_col(edgpu_ctx *c) {
  int algo = c->algo;
  if (algo == EDGPU_ALGO_AUTO)
    algo = (fast_supported_col(c, 0) && fast_supported_col(c, 1)) ? EDGPU_ALGO_FAST : (tiled_supported(c) ? EDGPU_ALGO_TILED : EDGPU_ALGO_GATHER);
  return algo;
}

bool hxv_fast_path(edgpu_ctx *c, const double *d_x) {
  if (c->orbs || c->dimph > 1 || c->dp.jhflag || c->opt_no_fuse) return false;
  if (c->nranks == 1) return pick_local(c) == EDGPU_ALGO_FAST && fast_supported_local(c);
  return (c->algo == EDGPU_ALGO_AUTO || c->algo == EDGPU_ALGO_FAST) && fast_peer_ready(c) && fast_supported_local(c);
}

// Phonon slabs (DimPh > 1): y(:, iph) += w0 iph x(:, iph) + E(i_el) [sqrt(iph+1) x(:, iph+1) + sqrt(iph) x(:, iph-1)] with
// E = sum_orb g_orb (n_up + n_dw - 1): spH0_ph, spH0e_eph x spH0ph_eph (stored/H_ph.f90, H_e_ph.f90;
// ED_HAMILTONIAN_SPARSE_HxV.f90:445-468), elementwise on top of the electronic H*v of every slab.
struct PhArgs { double g[EDGPU_MAX_ORB]; double w0; int norb, dimph; };
__global__ void k_phonon(PhArgs P, const int32_t *__restrict__ map_up, const int32_t *__restrict__ map_dw, int64_t dimup,
                         int64_t coloff, int64_t nel, const double *__restrict__ x, double *__restrict__ y) {
  for (int64_t ie = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; ie < nel; ie += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t mu = (uint32_t)map_up[ie % dimup], md = (uint32_t)map_dw[coloff + ie / dimup];
    double e = 0.0;
    for (int o = 0; o < P.norb; o++) e = e + P.g[o] * ((double)(((mu >> o) & 1u) + ((md >> o) & 1u)) - 1.0);
    for (int iph = 0; iph < P.dimph; iph++) {
      const int64_t i = ie + (int64_t)iph * nel;
      double acc = y[i] + (P.w0 * (double)iph) * x[i];
      if (iph + 1 < P.dimph) acc = acc + (e * sqrt((double)(iph + 1))) * x[i + nel];
      if (iph > 0) acc = acc + (e * sqrt((double)iph)) * x[i - nel];
      y[i] = acc;
    }
  }
}
static int hxv_apply_el(edgpu_ctx *c, const double *d_x, double *d_y);
int hxv_apply(edgpu_ctx *c, const double *d_x, double *d_y) {
  if (c->orbs) return orbs_apply(c, d_x, d_y);                     // ed_total_ud = F (spMatVec_orbs)
  if (c->dimph == 1) return hxv_apply_el(c, d_x, d_y);
  for (int iph = 0; iph < c->dimph; iph++) TRY(hxv_apply_el(c, d_x + (int64_t)iph * c->nel, d_y + (int64_t)iph * c->nel));
  PhArgs P{};
  for (int o = 0; o < EDGPU_MAX_ORB; o++) P.g[o] = c->hp.g_ph[o];
  P.w0 = c->hp.w0_ph; P.norb = c->dp.norb; P.dimph = c->dimph;
  prof_mark(c, "k_phonon");
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((c->nel + 255) / 256, (int64_t)c->sm_count * 8));
  k_phonon<<<grid, 256, 0, c->stream>>>(P, c->up.d_map, c->dw.d_map, c->dimup, c->coloff, c->nel, d_x, d_y);
  CKL(c);
  return EDGPU_OK;
}
// the electronic operator on ONE phonon slab (the whole vector when DimPh = 1)
static int hxv_apply_el(edgpu_ctx *c, const double *d_x, double *d_y) {
  if (c->nranks == 1) {
    const int algo = pick_local(c);
    if (algo == EDGPU_ALGO_FAST) {
      if (!fast_supported_local(c)) return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "fast H*v does not cover this sector/model");
      return fast_apply_local(c, d_x, d_y);
    }
    if (algo == EDGPU_ALGO_TILED) {
      if (!tiled_supported(c)) return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "tiled H*v does not cover this sector/model");
      return tiled_apply_local(c, d_x, d_y);
    }
    prof_mark(c, "k_hxv_gather");
    return gather_local(c, d_x, d_y, true, d_x);
  }
  // sharded, fast path: the same two kernels as on one GPU; i_dw sources that live on another rank arrive in the
  // local halo buffer, stored there by their owner over NVLink (no transposes, no packing)
  if ((c->algo == EDGPU_ALGO_AUTO || c->algo == EDGPU_ALGO_FAST) && !c->dp.jhflag && fast_peer_ready(c) &&
      fast_supported_local(c))
    return fast_apply_local(c, d_x, d_y);
  // sharded: diag + up locally; dw through the all-to-all transpose (spMatVec_MPI_main order,
  // ED_HAMILTONIAN_SPARSE_HxV.f90:587-644); non-local terms on the all-gathered vector (:673-692)
  const double *d_full = nullptr;
  if (c->dp.jhflag) {
    TRY(ensure(&c->d_full, c->dimup * c->dimdw));
    TRY(comm_allgather(c, d_x, c->d_full));
    d_full = c->d_full;
  }
  const int algo = pick_col(c);
  if (algo == EDGPU_ALGO_FAST && !(fast_supported_col(c, 0) && fast_supported_col(c, 1)))
    return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "fast H*v does not cover this sector/model");
  if (algo == EDGPU_ALGO_TILED && !tiled_supported(c))
    return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "tiled H*v does not cover this sector/model");
  if (algo == EDGPU_ALGO_FAST) TRY(fast_apply_col(c, 0, true, false, d_x, d_y, c->qdw, c->coloff));
  else if (algo == EDGPU_ALGO_TILED) TRY(tiled_apply_col(c, 0, true, d_x, d_y, c->qdw, c->coloff));
  else TRY(gather_local(c, d_x, d_y, false, d_full));
  TRY(ensure(&c->d_vt, c->dimdw * c->qup));
  TRY(ensure(&c->d_hvt, c->dimdw * c->qup));
  TRY(comm_transpose_fwd(c, d_x, c->d_vt));
  if (algo == EDGPU_ALGO_FAST) TRY(fast_apply_col(c, 1, false, false, c->d_vt, c->d_hvt, c->qup, c->rowoff));
  else if (algo == EDGPU_ALGO_TILED) TRY(tiled_apply_col(c, 1, false, c->d_vt, c->d_hvt, c->qup, c->rowoff));
  else TRY(gather_transposed_dw(c, c->d_vt, c->d_hvt));
  TRY(comm_transpose_bwd_add(c, c->d_hvt, d_y));
  return EDGPU_OK;
}
