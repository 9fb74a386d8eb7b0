// lanczos.cu -- device-resident Lanczos recurrences and Green's-function chains.
//
// sp_lanc_eigh / sp_lanc_tridiag are SciFortran routines (SF_SP_LINALG; un-vendored dependency of
// the reference, CMakeLists.txt:91-106) called at ED_DIAG.f90:174-186 and ED_GF_NORMAL.f90:232-237.
// Their recurrence per step is
//     k=1: v <- v/|v| ;  k>1: t <- v, v <- w/b, w <- -b t ;  w <- w + H v ; a = v.w ;
//     w <- w - a v ; b = |w|
// Here all vectors stay in HBM and normalisations are carried as scalar factors (v_k = sx * X), so
// a step is: H*x, one fused "w = sx*Hx - c*xp, a += (sx*x).w" pass and one fused
// "w -= a*sx*x, b2 += w.w" pass; the two scalars are reduced on device (NCCL all-reduce across
// ranks) and never visit the host inside a chain.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <complex>
#include <vector>

#include "engine.h"

#define RED_BLOCKS 1184          // 148 SMs x 8
#define RED_THREADS 256

__device__ __forceinline__ double block_sum(double v) {
  __shared__ double sh[RED_THREADS / 32];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x < 32) {
    r = (threadIdx.x < RED_THREADS / 32) ? sh[threadIdx.x] : 0.0;
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  __syncthreads();
  return r;                       // valid in thread 0
}

__global__ void __launch_bounds__(RED_THREADS) k_norm2(const double *__restrict__ x, int64_t n, double *__restrict__ partials) {
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s += x[i] * x[i];
  s = block_sum(s);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}
// deterministic fixed-order sum of the per-block partials -> st->red
__global__ void __launch_bounds__(RED_THREADS) k_finalize(const double *__restrict__ partials, int nb, LancState *st) {
  double s = 0.0;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) s += partials[i];
  s = block_sum(s);
  if (threadIdx.x == 0) st->red = s;
}
// w = sx*t - cprev*xp (in place over xp), partial of a = (sx*x).w
__global__ void __launch_bounds__(RED_THREADS) k_lanc_a(const double *__restrict__ t, const double *__restrict__ x,
                                                        double *__restrict__ xp, int64_t n, const LancState *st,
                                                        double *__restrict__ partials) {
  const double sx = st->sx, cp = st->cprev;
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double w = sx * t[i] - cp * xp[i];
    xp[i] = w;
    s += (sx * x[i]) * w;
  }
  s = block_sum(s);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}
// w -= a*(sx*x), partial of b^2 = w.w
__global__ void __launch_bounds__(RED_THREADS) k_lanc_b(double *__restrict__ w, const double *__restrict__ x, int64_t n,
                                                        const LancState *st, double *__restrict__ partials) {
  const double sx = st->sx, a = st->alpha;
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double v = w[i] - a * (sx * x[i]);
    w[i] = v;
    s += v * v;
  }
  s = block_sum(s);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}
__global__ void k_post_norm(LancState *st) { st->norm2 = st->red; st->sx = 1.0 / sqrt(st->red); st->cprev = 0.0; }
__global__ void k_post_alpha(LancState *st, double *alanc, int k) { st->alpha = st->red; alanc[k] = st->red; }
__global__ void k_post_beta(LancState *st, double *blanc, int k) {
  double b = sqrt(st->red);
  st->beta = b;
  blanc[k + 1] = b;               // blanc(k+1) = b_k ; blanc(1) = 0
  st->cprev = b * st->sx;         // next step: w = H v_{k+1} - b_k v_k , v_k = sx_old * X_old
  st->sx = 1.0 / b;               // v_{k+1} = w / b_k
}
// second sweep of sp_lanc_eigh (all scalars known): w = sx*t - cprev*xp - (a*sx)*x over xp,
// vect += (zk*sx)*x
__global__ void __launch_bounds__(RED_THREADS) k_lanc_sweep(const double *__restrict__ t, const double *__restrict__ x,
                                                            double *__restrict__ xp, double *__restrict__ vect, int64_t n,
                                                            double sx, double cprev, double a, double zk) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double xv = sx * x[i];
    xp[i] = (sx * t[i] - cprev * xp[i]) - a * xv;
    vect[i] += zk * xv;
  }
}
__global__ void k_scale(double *__restrict__ x, int64_t n, double s) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] *= s;
}
// start vector when the caller passes all zeros: counter-based hash of the GLOBAL index, so the
// vector does not depend on the rank count (the reference draws random_number with a fixed seed,
// which is compiler specific -- pass an explicit start vector for parity runs)
__global__ void k_random_start(double *__restrict__ x, int64_t n, int64_t goff) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t z = (uint64_t)(goff + i) + 0x9E3779B97F4A7C15ull * 1234567ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    x[i] = (double)(z >> 11) * (1.0 / 9007199254740992.0);
  }
}

static int ensure_coeffs(edgpu_ctx *c, int n) {
  if (c->lanc_cap >= n + 2) return EDGPU_OK;
  cudaFree(c->d_alanc); cudaFree(c->d_blanc);
  CK(cudaMalloc(&c->d_alanc, (size_t)(n + 2) * sizeof(double)));
  CK(cudaMalloc(&c->d_blanc, (size_t)(n + 2) * sizeof(double)));
  c->lanc_cap = n + 2;
  return EDGPU_OK;
}
static int reduce_to_state(edgpu_ctx *c, int nb = RED_BLOCKS) {   // partials -> st->red, all-reduced over ranks
  k_finalize<<<1, RED_THREADS, 0, c->stream>>>(c->d_partials, nb, c->d_st);
  CKL(c);
  return comm_allreduce_scalar(c, &c->d_st->red);
}

// one Lanczos step on device; on entry X = c->d_lx (scale st->sx), Xp = c->d_lp; on exit rotated
static int lanczos_step(edgpu_ctx *c, int k /*0-based*/) {
  if (hxv_fast_path(c, c->d_lx)) {
    // w = sx*(H x) - cprev*xp and the alpha partials come out of the column kernel's epilogue
    int nb = 0;
    TRY(fast_apply_local(c, c->d_lx, c->d_lt, c->d_lp, &nb));
    TRY(reduce_to_state(c, nb));
  } else {
    TRY(hxv_apply(c, c->d_lx, c->d_lt));
    k_lanc_a<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(c->d_lt, c->d_lx, c->d_lp, c->nloc, c->d_st, c->d_partials);
    CKL(c);
    TRY(reduce_to_state(c));
  }
  k_post_alpha<<<1, 1, 0, c->stream>>>(c->d_st, c->d_alanc, k);
  CKL(c);
  k_lanc_b<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(c->d_lp, c->d_lx, c->nloc, c->d_st, c->d_partials);
  CKL(c);
  TRY(reduce_to_state(c));
  k_post_beta<<<1, 1, 0, c->stream>>>(c->d_st, c->d_blanc, k);
  CKL(c);
  std::swap(c->d_lx, c->d_lp);    // X <- w (scale 1/b), Xp <- old X
  return EDGPU_OK;
}
static int lanczos_begin(edgpu_ctx *c, int ncoef) {
  TRY(vec_alloc(c, &c->d_lx, c->nloc));
  TRY(vec_alloc(c, &c->d_lp, c->nloc));
  TRY(vec_alloc(c, &c->d_lt, c->nloc));
  TRY(ensure_coeffs(c, ncoef));
  CK(cudaMemsetAsync(c->d_alanc, 0, (size_t)c->lanc_cap * sizeof(double), c->stream));
  CK(cudaMemsetAsync(c->d_blanc, 0, (size_t)c->lanc_cap * sizeof(double), c->stream));
  CK(cudaMemsetAsync(c->d_lp, 0, (size_t)c->nloc * sizeof(double), c->stream));
  return EDGPU_OK;
}
static int lanczos_norm_start(edgpu_ctx *c) {   // st->sx = 1/|X|, st->norm2 = |X|^2
  k_norm2<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(c->d_lx, c->nloc, c->d_partials);
  CKL(c);
  TRY(reduce_to_state(c));
  k_post_norm<<<1, 1, 0, c->stream>>>(c->d_st);
  CKL(c);
  return EDGPU_OK;
}

// ---- host-side tridiagonal QL (implicit shifts, EISPACK tql2 algorithm) -----------------------
static int tql2(int n, std::vector<double> &d, std::vector<double> &e, std::vector<double> &z) {
  if (n == 1) return 0;
  for (int i = 1; i < n; i++) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  double f = 0.0, tst1 = 0.0;
  for (int l = 0; l < n; l++) {
    int j = 0;
    double h = fabs(d[l]) + fabs(e[l]);
    if (tst1 < h) tst1 = h;
    int m;
    for (m = l; m < n; m++) {
      double tst2 = tst1 + fabs(e[m]);
      if (tst2 == tst1) break;
    }
    if (m != l) {
      double tst2;
      do {
        if (j++ == 60) return l + 1;
        int l1 = l + 1, l2 = l1 + 1;
        double g = d[l];
        double p = (d[l1] - g) / (2.0 * e[l]);
        double r = hypot(p, 1.0);
        double sr = (p >= 0.0) ? fabs(r) : -fabs(r);
        d[l] = e[l] / (p + sr);
        d[l1] = e[l] * (p + sr);
        double dl1 = d[l1];
        h = g - d[l];
        for (int i = l2; i < n; i++) d[i] -= h;
        f += h;
        p = d[m];
        double cc = 1.0, c2 = cc, c3 = cc, el1 = e[l1], s = 0.0, s2 = 0.0;
        for (int i = m - 1; i >= l; i--) {
          c3 = c2; c2 = cc; s2 = s;
          g = cc * e[i];
          h = cc * p;
          r = hypot(p, e[i]);
          e[i + 1] = s * r;
          s = e[i] / r;
          cc = p / r;
          p = cc * d[i] - s * g;
          d[i + 1] = h + s * (cc * g + s * d[i]);
          for (int k = 0; k < n; k++) {
            h = z[k + (size_t)n * (i + 1)];
            z[k + (size_t)n * (i + 1)] = s * z[k + (size_t)n * i] + cc * h;
            z[k + (size_t)n * i] = cc * z[k + (size_t)n * i] - s * h;
          }
        }
        p = -s * s2 * c3 * el1 * e[l] / dl1;
        e[l] = s * p;
        d[l] = cc * p;
        tst2 = tst1 + fabs(e[l]);
      } while (tst2 > tst1);
    }
    d[l] += f;
  }
  for (int ii = 1; ii < n; ii++) {
    int i = ii - 1, k = i;
    double p = d[i];
    for (int j = ii; j < n; j++)
      if (d[j] < p) { k = j; p = d[j]; }
    if (k != i) {
      d[k] = d[i]; d[i] = p;
      for (int j = 0; j < n; j++) std::swap(z[j + (size_t)n * i], z[j + (size_t)n * k]);
    }
  }
  return 0;
}
static int tridiag_eig(int n, const double *alanc, const double *blanc, std::vector<double> &diag, std::vector<double> &z) {
  diag.assign(alanc, alanc + n);
  std::vector<double> sub((size_t)n + 1, 0.0);
  for (int i = 1; i < n; i++) sub[i] = blanc[i];
  z.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) z[i + (size_t)n * i] = 1.0;
  if (tql2(n, diag, sub, z)) return edgpu_set_err(EDGPU_ERR_INVALID, "tql2 did not converge");
  return EDGPU_OK;
}

// ---- sp_lanc_eigh -----------------------------------------------------------------------------
extern "C" int edgpu_sp_lanc_eigh(edgpu_ctx *c, double *egs, double *vect, int64_t nloc, int nitermax,
                                  int iverbose, double threshold, int ncheck,
                                  int *nlanc_out, double *alanc_out, double *blanc_out) {
  if (!c || !c->hstatus) return edgpu_set_err(EDGPU_ERR_INVALID, "sp_lanc_eigh: Hsector NOT set");
  if (nloc != c->nloc) return edgpu_set_err(EDGPU_ERR_INVALID, "sp_lanc_eigh: size(vect) != vecDim");
  if (nitermax < 1) return edgpu_set_err(EDGPU_ERR_INVALID, "sp_lanc_eigh: Nitermax < 1");
  if (ncheck <= 0) ncheck = 10;
  CK(cudaSetDevice(c->device));
  TRY(lanczos_begin(c, nitermax));
  TRY(vec_alloc(c, &c->d_l0, c->nloc));
  TRY(vec_alloc(c, &c->d_lv, c->nloc));
  // start vector
  CK(cudaMemcpyAsync(c->d_lx, vect, (size_t)nloc * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  TRY(lanczos_norm_start(c));
  CK(cudaMemcpyAsync(c->h_pinned, &c->d_st->norm2, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (c->h_pinned[0] == 0.0) {
    k_random_start<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(c->d_lx, c->nloc, c->coloff * c->dimup);
    CKL(c);
    TRY(lanczos_norm_start(c));
  }
  CK(cudaMemcpyAsync(c->d_l0, c->d_lx, (size_t)nloc * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  CK(cudaMemcpyAsync(c->h_pinned, &c->d_st->norm2, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  const double norm0 = sqrt(c->h_pinned[0]);

  std::vector<double> alanc((size_t)nitermax + 2, 0.0), blanc((size_t)nitermax + 2, 0.0), diag, z;
  int nlanc = 0;
  double e_prev = 0.0, a_last = 0.0;
  *egs = 0.0;
  for (int iter = 1; iter <= nitermax; iter++) {
    TRY(lanczos_step(c, iter - 1));
    CK(cudaMemcpyAsync(c->h_pinned, &c->d_st->alpha, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    const double a_ = c->h_pinned[0], b_ = c->h_pinned[1];
    a_last = a_;
    if (fabs(b_) < threshold) break;
    nlanc++;
    alanc[iter - 1] = a_;
    blanc[iter] = b_;
    if (nlanc >= ncheck) {
      TRY(tridiag_eig(nlanc, alanc.data(), blanc.data(), diag, z));
      double diff = e_prev - diag[0];
      e_prev = diag[0];
      if (iverbose) fprintf(stderr, "edgpu lanczos iter %d E0 %.15g dE %.3e\n", iter, diag[0], diff);
      if (nlanc > ncheck && fabs(diff) <= threshold) break;
    }
  }
  if (nlanc == 0) { nlanc = 1; alanc[0] = a_last; }
  TRY(tridiag_eig(nlanc, alanc.data(), blanc.data(), diag, z));
  *egs = diag[0];
  // second sweep: vect = sum_k Z(k,1) v_k with the recorded coefficients (no reductions needed)
  CK(cudaMemcpyAsync(c->d_lx, c->d_l0, (size_t)nloc * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  CK(cudaMemsetAsync(c->d_lp, 0, (size_t)nloc * sizeof(double), c->stream));
  CK(cudaMemsetAsync(c->d_lv, 0, (size_t)nloc * sizeof(double), c->stream));
  double sx = 1.0 / norm0, cprev = 0.0;
  for (int k = 0; k < nlanc; k++) {
    TRY(hxv_apply(c, c->d_lx, c->d_lt));
    k_lanc_sweep<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(c->d_lt, c->d_lx, c->d_lp, c->d_lv, c->nloc, sx, cprev,
                                                             alanc[k], z[(size_t)k]);
    CKL(c);
    std::swap(c->d_lx, c->d_lp);
    const double b = blanc[k + 1];
    cprev = b * sx;
    sx = (b != 0.0) ? 1.0 / b : 0.0;
  }
  k_norm2<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(c->d_lv, c->nloc, c->d_partials);
  CKL(c);
  TRY(reduce_to_state(c));
  CK(cudaMemcpyAsync(c->h_pinned, &c->d_st->red, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  k_scale<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(c->d_lv, c->nloc, 1.0 / sqrt(c->h_pinned[0]));
  CKL(c);
  CK(cudaMemcpyAsync(vect, c->d_lv, (size_t)nloc * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (nlanc_out) *nlanc_out = nlanc;
  if (alanc_out) memcpy(alanc_out, alanc.data(), (size_t)nlanc * sizeof(double));
  if (blanc_out) memcpy(blanc_out, blanc.data(), (size_t)nlanc * sizeof(double));
  return EDGPU_OK;
}

// ---- sp_lanc_tridiag on a device-resident start vector in c->d_lx ------------------------------
static int tridiag_device(edgpu_ctx *c, int nlanc, double threshold, double *alanc, double *blanc) {
  TRY(lanczos_norm_start(c));
  for (int k = 0; k < nlanc; k++) TRY(lanczos_step(c, k));
  std::vector<double> a((size_t)nlanc + 2), b((size_t)nlanc + 2);
  CK(cudaMemcpyAsync(a.data(), c->d_alanc, (size_t)(nlanc + 1) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(b.data(), c->d_blanc, (size_t)(nlanc + 2) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  // reference bookkeeping: alanc(k)=a_k ; if |b_k|<threshold exit ; blanc(k+1)=b_k (k<nlanc)
  for (int k = 0; k < nlanc; k++) { alanc[k] = 0.0; blanc[k] = 0.0; }
  for (int k = 0; k < nlanc; k++) {
    alanc[k] = a[k];
    const double bk = b[k + 1];
    if (!(fabs(bk) >= threshold)) break;          // also stops on NaN
    if (k + 1 < nlanc) blanc[k + 1] = bk;
  }
  return EDGPU_OK;
}

extern "C" int edgpu_sp_lanc_tridiag(edgpu_ctx *c, const double *vin, int64_t nloc, double *alanc,
                                     double *blanc, int nlanc, double threshold) {
  if (!c || !c->hstatus) return edgpu_set_err(EDGPU_ERR_INVALID, "sp_lanc_tridiag: Hsector NOT set");
  if (nloc != c->nloc) return edgpu_set_err(EDGPU_ERR_INVALID, "sp_lanc_tridiag: size(vin) != vecDim");
  if (nlanc < 1) return edgpu_set_err(EDGPU_ERR_INVALID, "sp_lanc_tridiag: size(alanc) < 1");
  CK(cudaSetDevice(c->device));
  TRY(lanczos_begin(c, nlanc));
  CK(cudaMemcpyAsync(c->d_lx, vin, (size_t)nloc * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  return tridiag_device(c, nlanc, threshold, alanc, blanc);
}

extern "C" int edgpu_time_lanczos_device(edgpu_ctx *c, int64_t nloc, double *d_v0, int reps, double *ms_total) {
  if (!c || !c->hstatus) return edgpu_set_err(EDGPU_ERR_INVALID, "Hsector NOT set");
  if (nloc != c->nloc) return edgpu_set_err(EDGPU_ERR_INVALID, "nloc mismatch");
  CK(cudaSetDevice(c->device));
  TRY(lanczos_begin(c, reps));
  CK(cudaMemcpyAsync(c->d_lx, d_v0, (size_t)nloc * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  TRY(lanczos_norm_start(c));
  CK(cudaEventRecord(c->ev0, c->stream));
  for (int k = 0; k < reps; k++) TRY(lanczos_step(c, k));
  CK(cudaEventRecord(c->ev1, c->stream));
  CK(cudaEventSynchronize(c->ev1));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  *ms_total = ms;
  return EDGPU_OK;
}

// ---- Green's function chains ------------------------------------------------------------------------
extern "C" int edgpu_gf_set_state(edgpu_ctx *c, int isector, const double *gs, int64_t nloc, double e0) {
  if (!c) return edgpu_set_err(EDGPU_ERR_INVALID, "ctx == NULL");
  CK(cudaSetDevice(c->device));
  int nup, ndw;
  TRY(edgpu_get_nup_ndw(c, isector, &nup, &ndw));
  int64_t vd;
  TRY(edgpu_vecdim_hv_sector(c, isector, &vd));
  if (vd != nloc) return edgpu_set_err(EDGPU_ERR_INVALID, "gf_set_state: nloc != vecDim(isector)");
  cudaFree(c->d_gs); c->d_gs = nullptr;
  CK(cudaMalloc(&c->d_gs, (size_t)nloc * sizeof(double)));
  CK(cudaMemcpyAsync(c->d_gs, gs, (size_t)nloc * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->gs_nup = nup; c->gs_ndw = ndw; c->gs_nloc = nloc; c->gs_e0 = e0;
  return EDGPU_OK;
}

// vvinit(j) = sgn * gs(i), |j> = c^+_{iorb,ispin}|i> or c_{iorb,ispin}|i>  (ED_GF_NORMAL.f90:190-209,
// 264-283), evaluated from the TARGET element: the source word is the target with the orbital's
// bit flipped back, its position the closed-form rank (no gathered vector, no master-only loop).
__global__ void k_gf_start(const int32_t *__restrict__ tmap_up, const int32_t *__restrict__ tmap_dw, int64_t tdimup,
                           int64_t tqdw, int64_t tcoloff, int64_t sdimup, int64_t scoloff,
                           const double *__restrict__ gs, int iorb, int ispin, int add,
                           const uint32_t *__restrict__ binom, double *__restrict__ out) {
  const uint32_t bit = 1u << (iorb - 1);
  for (int64_t jl = blockIdx.y; jl < tqdw; jl += gridDim.y) {
    int64_t scol = jl;                                    // ispin==1: same local column
    double csgn = 1.0;
    bool col_ok = true;
    if (ispin == 2) {
      uint32_t r = (uint32_t)tmap_dw[tcoloff + jl];
      bool has = (r & bit) != 0;
      col_ok = add ? has : !has;
      uint32_t m = add ? (r & ~bit) : (r | bit);
      csgn = hd_sign_below(m, iorb);
      scol = hd_rank(m, binom) - scoloff;                 // nranks==1 only: scoloff = 0
    }
    for (int64_t ju = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; ju < tdimup; ju += (int64_t)gridDim.x * blockDim.x) {
      double v = 0.0;
      if (ispin == 1) {
        uint32_t r = (uint32_t)tmap_up[ju];
        bool has = (r & bit) != 0;
        if (add ? has : !has) {
          uint32_t m = add ? (r & ~bit) : (r | bit);
          v = hd_sign_below(m, iorb) * gs[hd_rank(m, binom) + scol * sdimup];
        }
      } else if (col_ok) {
        v = csgn * gs[ju + scol * sdimup];
      }
      out[ju + jl * tdimup] = v;
    }
  }
}

extern "C" int edgpu_gf_chains(edgpu_ctx *c, int nchains, const int *iorb, const int *ispin,
                               const int *addrem, int nlanc_max, double threshold,
                               double *norm2, int *nlanc, double *alanc, double *blanc) {
  if (!c || !c->d_gs) return edgpu_set_err(EDGPU_ERR_INVALID, "gf_chains: no state set (edgpu_gf_set_state)");
  if (c->hstatus) return edgpu_set_err(EDGPU_ERR_INVALID, "gf_chains: a sector is live; call delete_Hv_sector first");
  CK(cudaSetDevice(c->device));
  std::vector<int> done((size_t)nchains, 0);
  for (int ch = 0; ch < nchains; ch++) {
    norm2[ch] = 0.0; nlanc[ch] = 0;
    for (int k = 0; k < nlanc_max; k++) { alanc[(size_t)ch * nlanc_max + k] = 0.0; blanc[(size_t)ch * nlanc_max + k] = 0.0; }
    if (iorb[ch] < 1 || iorb[ch] > c->dp.norb || ispin[ch] < 1 || ispin[ch] > 2 || (addrem[ch] != 1 && addrem[ch] != -1))
      return edgpu_set_err(EDGPU_ERR_INVALID, "gf_chains: bad channel %d", ch);
    if (ispin[ch] == 2 && c->nranks > 1)
      return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "gf_chains: spin-down operators on a sharded state need a column exchange (next row of SURVEY 8f)");
  }
  for (int ch = 0; ch < nchains; ch++) {
    if (done[ch]) continue;
    // target sector of this channel (getCDGsector / getCsector, ED_SETUP.f90:377-418)
    int jnup = c->gs_nup + (ispin[ch] == 1 ? addrem[ch] : 0);
    int jndw = c->gs_ndw + (ispin[ch] == 2 ? addrem[ch] : 0);
    if (jnup < 0 || jnup > c->ns || jndw < 0 || jndw > c->ns) { done[ch] = 1; continue; }   // jsector == 0
    int jsector;
    TRY(edgpu_get_sector(c, jnup, jndw, &jsector));
    TRY(edgpu_build_hv_sector(c, jsector));
    int rc = EDGPU_OK;
    for (int ch2 = ch; ch2 < nchains && !rc; ch2++) {       // every channel sharing this target sector
      if (done[ch2]) continue;
      int n2 = c->gs_nup + (ispin[ch2] == 1 ? addrem[ch2] : 0), d2 = c->gs_ndw + (ispin[ch2] == 2 ? addrem[ch2] : 0);
      if (n2 != jnup || d2 != jndw) continue;
      done[ch2] = 1;
      const int64_t jdim = c->dimup * c->dimdw;
      const int nl = (int)std::min<int64_t>(jdim, nlanc_max);   // nlanc=min(jdim,lanc_nGFiter), :219
      rc = lanczos_begin(c, nl);
      if (rc) break;
      const int64_t sdimup = c->h_binom[c->ns * EDGPU_BINOM_LD + c->gs_nup];
      dim3 grid((unsigned)((c->dimup + 255) / 256), (unsigned)(c->qdw < 32768 ? c->qdw : 32768));
      k_gf_start<<<grid, 256, 0, c->stream>>>(c->up.d_map, c->dw.d_map, c->dimup, c->qdw, c->coloff, sdimup, 0,
                                              c->d_gs, iorb[ch2], ispin[ch2], addrem[ch2] == 1 ? 1 : 0, c->d_binom, c->d_lx);
      c->launches++;
      if (cudaGetLastError() != cudaSuccess) { rc = edgpu_set_err(EDGPU_ERR_CUDA, "k_gf_start launch failed"); break; }
      rc = tridiag_device(c, nl, threshold, alanc + (size_t)ch2 * nlanc_max, blanc + (size_t)ch2 * nlanc_max);
      if (rc) break;
      if (cudaMemcpy(&norm2[ch2], &c->d_st->norm2, sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) {
        rc = edgpu_set_err(EDGPU_ERR_CUDA, "norm2 read-back failed");
        break;
      }
      nlanc[ch2] = nl;
    }
    edgpu_delete_hv_sector(c);
    if (rc) return rc;
  }
  return EDGPU_OK;
}

extern "C" int edgpu_add_to_lanczos_gf(double norm2, double zeta, double ei, const double *alanc,
                                       const double *blanc, int nlanc, int isign,
                                       const double *zin, int nz, double *gout) {
  if (nlanc < 1) return edgpu_set_err(EDGPU_ERR_INVALID, "add_to_lanczos_gf: nlanc < 1");
  std::vector<double> diag, z;
  TRY(tridiag_eig(nlanc, alanc, blanc, diag, z));
  const std::complex<double> *zz = reinterpret_cast<const std::complex<double> *>(zin);
  std::complex<double> *g = reinterpret_cast<std::complex<double> *>(gout);
  const double pesobz = norm2 / zeta;                       // T=0 branch, ED_GF_NORMAL.f90:615-621
  for (int j = 0; j < nlanc; j++) {
    const double de = diag[j] - ei;
    const double z1 = z[0 + (size_t)nlanc * j];
    const double peso = pesobz * z1 * z1;
    for (int i = 0; i < nz; i++) g[i] += peso / (zz[i] - (double)isign * de);
  }
  return EDGPU_OK;
}
