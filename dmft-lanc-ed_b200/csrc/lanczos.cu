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
#include <functional>
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
__global__ void __launch_bounds__(RED_THREADS) k_dot(const double *__restrict__ x, const double *__restrict__ y, int64_t n, double *__restrict__ partials) {
  double s = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s += x[i] * y[i];
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
// w -= a*(sx*x), partial of b^2 = w.w.  Pure streaming (24 B/element): 16-byte accesses, four independent pairs per
// thread and trip in flight, streaming cache hints (nothing is reused before the next pass evicts it).
__global__ void __launch_bounds__(RED_THREADS) k_lanc_b(double *__restrict__ w, const double *__restrict__ x, int64_t n,
                                                        const LancState *st, double *__restrict__ partials) {
  const double al = st->alpha, sx = st->sx;                        // same expression order as before: w - a*(sx*x)
  double s = 0.0;
  const int64_t n2 = n >> 1, stride = (int64_t)gridDim.x * blockDim.x;
  double2 *w2 = reinterpret_cast<double2 *>(w);
  const double2 *x2 = reinterpret_cast<const double2 *>(x);
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n2; i += 4 * stride) {
    double2 a[4], b[4];
#pragma unroll
    for (int q = 0; q < 4; q++) { a[q] = __ldcs(w2 + i + q * stride); b[q] = __ldcs(x2 + i + q * stride); }
#pragma unroll
    for (int q = 0; q < 4; q++) {
      a[q].x = a[q].x - al * (sx * b[q].x); a[q].y = a[q].y - al * (sx * b[q].y);
      __stcs(w2 + i + q * stride, a[q]);
      s += a[q].x * a[q].x; s += a[q].y * a[q].y;
    }
  }
  for (; i < n2; i += stride) {
    double2 a = __ldcs(w2 + i);
    const double2 b = __ldcs(x2 + i);
    a.x = a.x - al * (sx * b.x); a.y = a.y - al * (sx * b.y);
    __stcs(w2 + i, a);
    s += a.x * a.x; s += a.y * a.y;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {             // odd length: the last element
    const double v = w[n - 1] - al * (sx * x[n - 1]);
    w[n - 1] = v;
    s += v * v;
  }
  s = block_sum(s);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}
// what a reduction result means for the recurrence: KIND 0 = |x|^2 of the start vector, 1 = a_k, 2 = b_k^2
template <int KIND>
__device__ __forceinline__ void post_scalar(LancState *st, double *coef, int k) {
  if (KIND == 0) { st->norm2 = st->red; st->sx = 1.0 / sqrt(st->red); st->cprev = 0.0; }
  if (KIND == 1) { st->alpha = st->red; coef[k] = st->red; }
  if (KIND == 2) {
    const double b = sqrt(st->red);
    st->beta = b;
    coef[k + 1] = b;              // blanc(k+1) = b_k ; blanc(1) = 0
    st->cprev = b * st->sx;       // next step: w = H v_{k+1} - b_k v_k , v_k = sx_old * X_old
    st->sx = 1.0 / b;             // v_{k+1} = w / b_k
  }
}
template <int KIND>
__global__ void k_post(LancState *st, double *coef, int k) { post_scalar<KIND>(st, coef, k); }
// single rank: the final reduction and its interpretation in one launch
template <int KIND>
__global__ void __launch_bounds__(RED_THREADS) k_finalize_post(const double *__restrict__ partials, int nb, LancState *st, double *coef, int k) {
  double s = 0.0;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) s += partials[i];
  s = block_sum(s);
  if (threadIdx.x == 0) { st->red = s; post_scalar<KIND>(st, coef, k); }
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
// scalars of one step of the second sweep into the device state (read by k_fcol's Lanczos epilogue)
__global__ void k_set_sweep(LancState *st, double sx, double cprev, double a, double zk) {
  st->sx = sx; st->cprev = cprev; st->sw_a = a; st->sw_zk = zk;
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
// partials -> st->red (all ranks) -> Lanczos scalar of kind KIND; one launch on a single rank
template <int KIND>
static int reduce_and_post(edgpu_ctx *c, int nb, double *coef, int k) {
  if (c->nranks == 1) {
    k_finalize_post<KIND><<<1, RED_THREADS, 0, c->stream>>>(c->d_partials, nb, c->d_st, coef, k);
    CKL(c);
    return EDGPU_OK;
  }
  TRY(reduce_to_state(c, nb));
  k_post<KIND><<<1, 1, 0, c->stream>>>(c->d_st, coef, k);
  CKL(c);
  return EDGPU_OK;
}

// one Lanczos step on device; on entry X = c->d_lx (scale st->sx), Xp = c->d_lp; on exit rotated
static int lanczos_step(edgpu_ctx *c, int k /*0-based*/) {
  if (hxv_fast_path(c, c->d_lx)) {
    // w = sx*(H x) - cprev*xp and the alpha partials come out of the column kernel's epilogue
    int nb = 0;
    TRY(fast_apply_local(c, c->d_lx, c->d_lt, c->d_lp, &nb));
    TRY(reduce_and_post<1>(c, nb, c->d_alanc, k));
  } else {
    TRY(hxv_apply(c, c->d_lx, c->d_lt));
    k_lanc_a<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(c->d_lt, c->d_lx, c->d_lp, c->nloc, c->d_st, c->d_partials);
    CKL(c);
    TRY(reduce_and_post<1>(c, RED_BLOCKS, c->d_alanc, k));
  }
  k_lanc_b<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(c->d_lp, c->d_lx, c->nloc, c->d_st, c->d_partials);
  CKL(c);
  TRY(reduce_and_post<2>(c, RED_BLOCKS, c->d_blanc, k));
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
  return reduce_and_post<0>(c, RED_BLOCKS, nullptr, 0);
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

// Lowest eigenvalue of the n x n tridiagonal (diagonal a[0..n), sub-diagonal b[1..n)) by Sturm-sequence bisection:
// O(n) per probe, no eigenvectors.  Used for the running convergence test of sp_lanc_eigh (the reference re-runs a
// full tql2 with eigenvectors at every step, O(n^3) each); an algorithm independent of tql2 above.
static double tridiag_lowest(int n, const double *a, const double *b) {
  double lo = a[0], hi = a[0];
  for (int i = 0; i < n; i++) {                                    // Gershgorin interval
    const double r = (i > 0 ? fabs(b[i]) : 0.0) + (i + 1 < n ? fabs(b[i + 1]) : 0.0);
    lo = std::min(lo, a[i] - r);
    hi = std::max(hi, a[i] + r);
  }
  if (n == 1) return a[0];
  const double tiny = 1e-300;
  auto below = [&](double x) {                                     // number of eigenvalues < x
    int cnt = 0;
    double q = a[0] - x;
    if (q < 0.0) cnt++;
    for (int i = 1; i < n; i++) {
      if (fabs(q) < tiny) q = (q < 0.0) ? -tiny : tiny;
      q = a[i] - x - b[i] * b[i] / q;
      if (q < 0.0) cnt++;
    }
    return cnt;
  };
  for (int it = 0; it < 200; it++) {
    const double mid = 0.5 * (lo + hi);
    if (mid <= lo || mid >= hi) break;                             // interval is one ulp wide
    if (below(mid) >= 1) hi = mid; else lo = mid;
  }
  return 0.5 * (lo + hi);
}

// ---- sp_lanc_eigh -----------------------------------------------------------------------------
// Core of sp_lanc_eigh on device vectors: on entry c->d_lx holds the start vector (all zero => pseudo-random),
// on exit c->d_lv the normalised eigenvector; nothing but scalars crosses the host boundary.
static int lanc_eigh_core(edgpu_ctx *c, int nitermax, int iverbose, double threshold, int ncheck, double *egs,
                          int *nlanc_out, std::vector<double> &alanc, std::vector<double> &blanc) {
  const int64_t nloc = c->nloc;
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

  alanc.assign((size_t)nitermax + 2, 0.0); blanc.assign((size_t)nitermax + 2, 0.0);
  std::vector<double> diag, z;
  int nlanc = 0;
  double e_prev = 0.0, a_last = 0.0;
  *egs = 0.0;
  // The recurrence runs on the device in blocks of `ncheck` steps; the host looks at a block's coefficients while
  // the NEXT block is already running (one event wait per block, never per step), and applies the reference's
  // per-step tests to them: exit on |b_k| < threshold, or on |E0(k) - E0(k-1)| <= threshold once ncheck steps are in.
  // Steps the device ran past the stopping point are simply not used (the second sweep replays the accepted ones).
  const size_t cap = (size_t)nitermax + 2;
  double *h_coef = nullptr;                                        // [2 blocks in flight][alanc | blanc]
  CK(cudaMallocHost(&h_coef, 4 * cap * sizeof(double)));
  cudaEvent_t evb[2] = {nullptr, nullptr};
  int rc = EDGPU_OK;
  for (int q = 0; q < 2 && !rc; q++) if (cudaEventCreateWithFlags(&evb[q], cudaEventDisableTiming) != cudaSuccess) rc = edgpu_set_err(EDGPU_ERR_CUDA, "event");
  int enq = 0, seen = 0, nblocks = 0;
  bool stop = false;
  auto enqueue_block = [&]() -> int {
    const int bs = std::min(ncheck, nitermax - enq);
    for (int k = 0; k < bs; k++) TRY(lanczos_step(c, enq + k));
    enq += bs;
    double *h = h_coef + (size_t)(nblocks & 1) * 2 * cap;
    CK(cudaMemcpyAsync(h, c->d_alanc, (size_t)enq * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(h + cap, c->d_blanc, (size_t)(enq + 1) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaEventRecord(evb[nblocks & 1], c->stream));
    nblocks++;
    return EDGPU_OK;
  };
  auto examine_block = [&](int blk, int upto) -> int {            // steps [seen, upto) of block blk have landed
    CK(cudaEventSynchronize(evb[blk & 1]));
    const double *h = h_coef + (size_t)(blk & 1) * 2 * cap;
    for (int k = seen; k < upto && !stop; k++) {
      const double a_ = h[k], b_ = h[cap + k + 1];
      a_last = a_;
      if (!(fabs(b_) >= threshold)) { stop = true; break; }        // also stops on NaN
      nlanc++;
      alanc[(size_t)k] = a_;
      blanc[(size_t)k + 1] = b_;
      if (nlanc >= ncheck) {
        const double e0 = tridiag_lowest(nlanc, alanc.data(), blanc.data());
        const double diff = e_prev - e0;
        e_prev = e0;
        if (iverbose) fprintf(stderr, "edgpu lanczos iter %d E0 %.15g dE %.3e\n", k + 1, e0, diff);
        if (nlanc > ncheck && fabs(diff) <= threshold) stop = true;
      }
    }
    seen = upto;
    return EDGPU_OK;
  };
  if (!rc) rc = enqueue_block();
  while (!rc && !stop) {
    const int prev_blk = nblocks - 1, prev_upto = enq;
    if (enq < nitermax) rc = enqueue_block();                      // the device keeps working ...
    if (!rc) rc = examine_block(prev_blk, prev_upto);              // ... while the host tests the block before
    if (!rc && !stop && seen == nitermax) break;
  }
  cudaStreamSynchronize(c->stream);
  for (int q = 0; q < 2; q++) if (evb[q]) cudaEventDestroy(evb[q]);
  cudaFreeHost(h_coef);
  if (rc) return rc;
  if (nlanc == 0) { nlanc = 1; alanc[0] = a_last; }
  TRY(tridiag_eig(nlanc, alanc.data(), blanc.data(), diag, z));
  *egs = diag[0];
  // second sweep: vect = sum_k Z(k,1) v_k with the recorded coefficients (no reductions needed)
  CK(cudaMemcpyAsync(c->d_lx, c->d_l0, (size_t)nloc * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  CK(cudaMemsetAsync(c->d_lp, 0, (size_t)nloc * sizeof(double), c->stream));
  CK(cudaMemsetAsync(c->d_lv, 0, (size_t)nloc * sizeof(double), c->stream));
  double sx = 1.0 / norm0, cprev = 0.0;
  const bool fused_sweep = hxv_fast_path(c, c->d_lx);
  for (int k = 0; k < nlanc; k++) {
    if (fused_sweep) {
      // the whole vector update of the step rides in the column kernel's epilogue: w = sx*Hx - cprev*xp - a*v over xp,
      // vect += Z(k,1)*v (ED_DIAG.f90:174-186 -> sp_lanc_eigh's second recurrence)
      k_set_sweep<<<1, 1, 0, c->stream>>>(c->d_st, sx, cprev, alanc[k], z[(size_t)k]);
      CKL(c);
      c->sweep_vect = c->d_lv;
      int nb = 0;
      const int rc2 = fast_apply_local(c, c->d_lx, c->d_lt, c->d_lp, &nb);
      c->sweep_vect = nullptr;
      if (rc2) return rc2;
    } else {
      TRY(hxv_apply(c, c->d_lx, c->d_lt));
      k_lanc_sweep<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(c->d_lt, c->d_lx, c->d_lp, c->d_lv, c->nloc, sx, cprev,
                                                               alanc[k], z[(size_t)k]);
      CKL(c);
    }
    std::swap(c->d_lx, c->d_lp);
    const double b = blanc[k + 1];
    cprev = b * sx;
    sx = (b != 0.0) ? 1.0 / b : 0.0;
  }
  if (fused_sweep) {                                               // leave no sweep scalars behind for the next recurrence
    k_set_sweep<<<1, 1, 0, c->stream>>>(c->d_st, 0.0, 0.0, 0.0, 0.0);
    CKL(c);
  }
  k_norm2<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(c->d_lv, c->nloc, c->d_partials);
  CKL(c);
  TRY(reduce_to_state(c));
  CK(cudaMemcpyAsync(c->h_pinned, &c->d_st->red, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  k_scale<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(c->d_lv, c->nloc, 1.0 / sqrt(c->h_pinned[0]));
  CKL(c);
  c->lv_valid = true;
  c->gs_e0 = *egs;
  if (nlanc_out) *nlanc_out = nlanc;
  return EDGPU_OK;
}

extern "C" int edgpu_sp_lanc_eigh(edgpu_ctx *c, double *egs, double *vect, int64_t nloc, int nitermax,
                                  int iverbose, double threshold, int ncheck,
                                  int *nlanc_out, double *alanc_out, double *blanc_out) {
  if (!c || !c->hstatus) return edgpu_set_err(EDGPU_ERR_INVALID, "sp_lanc_eigh: Hsector NOT set");
  if (nloc != c->nloc) return edgpu_set_err(EDGPU_ERR_INVALID, "sp_lanc_eigh: size(vect) != vecDim");
  if (nitermax < 1) return edgpu_set_err(EDGPU_ERR_INVALID, "sp_lanc_eigh: Nitermax < 1");
  if (ncheck <= 0) ncheck = 10;
  CK(cudaSetDevice(c->device));
  c->lv_valid = false;
  TRY(lanczos_begin(c, nitermax));
  TRY(vec_alloc(c, &c->d_l0, c->nloc));
  TRY(vec_alloc(c, &c->d_lv, c->nloc));
  CK(cudaMemcpyAsync(c->d_lx, vect, (size_t)nloc * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  std::vector<double> alanc, blanc;
  int nlanc = 0;
  TRY(lanc_eigh_core(c, nitermax, iverbose, threshold, ncheck, egs, &nlanc, alanc, blanc));
  CK(cudaMemcpyAsync(vect, c->d_lv, (size_t)nloc * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (nlanc_out) *nlanc_out = nlanc;
  if (alanc_out) memcpy(alanc_out, alanc.data(), (size_t)nlanc * sizeof(double));
  if (blanc_out) memcpy(blanc_out, blanc.data(), (size_t)nlanc * sizeof(double));
  return EDGPU_OK;
}

// ---- sector scan (ed_diag_d, ED_DIAG.f90:83-276) ---------------------------------------------------------------
// For every listed sector: build_Hv_sector, lowest eigenpair by Lanczos from the pseudo-random start vector (the
// reference: dense eigh below lanc_dim_threshold, sp_lanc_eigh above; Nitermax = min(dim, nitermax)), delete.  Twin
// sectors (Nup <-> Ndw, ed_twin: same spectrum when the two spins have the same parameters) are diagonalised once.
// The eigenvector of the lowest sector stays on the device as the state of the GF chains / observables
// (edgpu_gf_set_state_from_eigh semantics), so nothing but the energies returns to the host.
extern "C" int edgpu_diag_sectors(edgpu_ctx *c, int nsectors, const int *isector, int nitermax, double threshold,
                                  int ncheck, int twin, double *e0, int *nlanc, int *best) {
  if (!c || !isector || !e0 || nsectors < 1) return edgpu_set_err(EDGPU_ERR_INVALID, "diag_sectors: bad arguments");
  if (c->hstatus) return edgpu_set_err(EDGPU_ERR_INVALID, "diag_sectors: a sector is live; call delete_Hv_sector first");
  if (ncheck <= 0) ncheck = 10;
  CK(cudaSetDevice(c->device));
  int ibest = -1;
  std::vector<int> src((size_t)nsectors, -1);                      // twin: index of the sector whose result is copied
  for (int s = 0; s < nsectors && c->hp.ed_total_ud; s++) {        // (twin reuse for ed_total_ud = F: not built, every sector is solved)
    int nup, ndw;
    TRY(edgpu_get_nup_ndw(c, isector[s], &nup, &ndw));
    if (twin && nup < ndw)
      for (int t = 0; t < nsectors; t++) {
        int nu2, nd2;
        TRY(edgpu_get_nup_ndw(c, isector[t], &nu2, &nd2));
        if (nu2 == ndw && nd2 == nup) { src[(size_t)s] = t; break; }
      }
  }
  for (int s = 0; s < nsectors; s++) {
    if (src[(size_t)s] >= 0) continue;
    TRY(edgpu_build_hv_sector(c, isector[s]));
    const int64_t dim = c->dimup * c->dimdw;
    const int nit = (int)std::min<int64_t>(dim, nitermax);
    int rc = lanczos_begin(c, nit);
    if (!rc) rc = vec_alloc(c, &c->d_l0, c->nloc);
    if (!rc) rc = vec_alloc(c, &c->d_lv, c->nloc);
    if (!rc && cudaMemsetAsync(c->d_lx, 0, (size_t)std::max<int64_t>(c->nloc, 1) * sizeof(double), c->stream) != cudaSuccess) rc = edgpu_set_err(EDGPU_ERR_CUDA, "diag_sectors: memset");
    std::vector<double> al, bl;
    int nl = 0;
    if (!rc) rc = lanc_eigh_core(c, nit, 0, threshold, ncheck, &e0[s], &nl, al, bl);
    if (nlanc) nlanc[s] = nl;
    if (!rc && (ibest < 0 || e0[s] < e0[ibest])) {
      ibest = s;
      if (c->hp.ed_total_ud && c->dimph == 1) rc = edgpu_gf_set_state_from_eigh(c);   // (no chains / observables for ed_total_ud = F or DimPh > 1)
    }
    edgpu_delete_hv_sector(c);
    if (rc) return rc;
  }
  for (int s = 0; s < nsectors; s++)
    if (src[(size_t)s] >= 0) { e0[s] = e0[src[(size_t)s]]; if (nlanc) nlanc[s] = nlanc[src[(size_t)s]]; }
  if (best) *best = ibest;
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

// <a, b> over the whole sector vector (local shards, all-reduced over the ranks): the reference's
// dot_product + MPI_Allreduce of its MPI Lanczos, exposed for hosts that keep vectors in HBM.  Collective.
extern "C" int edgpu_dev_dot(edgpu_ctx *c, int64_t nloc, const double *d_a, const double *d_b, double *out) {
  if (!c || !c->hstatus) return edgpu_set_err(EDGPU_ERR_INVALID, "dev_dot: Hsector NOT set");
  if (nloc != c->nloc) return edgpu_set_err(EDGPU_ERR_INVALID, "dev_dot: nloc != vecDim");
  CK(cudaSetDevice(c->device));
  k_dot<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(d_a, d_b, nloc, c->d_partials);
  CKL(c);
  TRY(reduce_to_state(c));
  CK(cudaMemcpyAsync(c->h_pinned, &c->d_st->red, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  *out = c->h_pinned[0];
  return EDGPU_OK;
}

// ---- Green's function chains ------------------------------------------------------------------------
extern "C" int edgpu_gf_set_state(edgpu_ctx *c, int isector, const double *gs, int64_t nloc, double e0) {
  if (!c) return edgpu_set_err(EDGPU_ERR_INVALID, "ctx == NULL");
  if (!c->hp.ed_total_ud) return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "ed_total_ud = F: the chains / observables of an orbital-resolved state are not built");
  if (c->dimph > 1) return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "DimPh > 1: the chains / observables of an electron-phonon state are not built");
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

// The same hand-off without the host: the eigenvector the last edgpu_sp_lanc_eigh left on the device (sector still
// live) becomes the state of the chains.  Replaces es_return_cvector's gather to the master rank
// (ED_EIGENSPACE.f90:502-572): every rank keeps its own shard.
extern "C" int edgpu_gf_set_state_from_eigh(edgpu_ctx *c) {
  if (!c || !c->hstatus) return edgpu_set_err(EDGPU_ERR_INVALID, "gf_set_state_from_eigh: Hsector NOT set");
  if (c->orbs || c->dimph > 1) return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "chains / observables of an orbital-resolved or electron-phonon state are not built");
  if (!c->lv_valid || !c->d_lv) return edgpu_set_err(EDGPU_ERR_INVALID, "gf_set_state_from_eigh: no eigenvector of this sector on the device (call edgpu_sp_lanc_eigh first)");
  CK(cudaSetDevice(c->device));
  cudaFree(c->d_gs); c->d_gs = nullptr;
  CK(cudaMalloc(&c->d_gs, (size_t)std::max<int64_t>(c->nloc, 1) * sizeof(double)));
  CK(cudaMemcpyAsync(c->d_gs, c->d_lv, (size_t)c->nloc * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->gs_nup = c->nup; c->gs_ndw = c->ndw; c->gs_nloc = c->nloc;
  return EDGPU_OK;
}

// vvinit(j) = sgn * gs(i), |j> = c^+_{iorb,ispin}|i> or c_{iorb,ispin}|i>  (ED_GF_NORMAL.f90:190-209,
// 264-283), evaluated from the TARGET element: the source word is the target with the orbital's
// bit flipped back, its position the closed-form rank (no gathered vector, no master-only loop).
__global__ void k_gf_start(const int32_t *__restrict__ tmap_up, const int32_t *__restrict__ tmap_dw, int64_t tdimup,
                           int64_t tqdw, int64_t tcoloff, int64_t sdimup,
                           const double *__restrict__ gs, int iorb, int add,
                           const uint32_t *__restrict__ binom, double *__restrict__ out) {
  // spin-up operator: acts inside a column (the local dw columns of the target and of the state coincide)
  const uint32_t bit = 1u << (iorb - 1);
  (void)tmap_dw; (void)tcoloff;
  for (int64_t jl = blockIdx.y; jl < tqdw; jl += gridDim.y) {
    for (int64_t ju = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; ju < tdimup; ju += (int64_t)gridDim.x * blockDim.x) {
      double v = 0.0;
      const uint32_t r = (uint32_t)tmap_up[ju];
      const bool has = (r & bit) != 0;
      if (add ? has : !has) {
        const uint32_t m = add ? (r & ~bit) : (r | bit);
        v = hd_sign_below(m, iorb) * gs[hd_rank(m, binom) + jl * sdimup];
      }
      out[ju + jl * tdimup] = v;
    }
  }
}

// spin-down operator: target column jl is sgn[jl] times column src[jl] of S (src < 0: zero column)
__global__ void k_gf_start_dw(const double *__restrict__ S, const int *__restrict__ src, const double *__restrict__ sgn,
                              int64_t dimup, int64_t tqdw, double *__restrict__ out) {
  for (int64_t jl = blockIdx.y; jl < tqdw; jl += gridDim.y) {
    const int sc = src[jl];
    const double sg = sgn[jl];
    for (int64_t ju = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; ju < dimup; ju += (int64_t)gridDim.x * blockDim.x)
      out[ju + jl * dimup] = sc >= 0 ? sg * S[ju + (int64_t)sc * dimup] : 0.0;
  }
}
// gathers whole columns: out(:, i) = in(:, idx[i])
__global__ void k_gather_cols(const double *__restrict__ in, const int *__restrict__ idx, int64_t dimup, int64_t ncols, double *__restrict__ out) {
  for (int64_t i = blockIdx.y; i < ncols; i += gridDim.y) {
    const int64_t sc = idx[i];
    for (int64_t ju = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; ju < dimup; ju += (int64_t)gridDim.x * blockDim.x)
      out[ju + i * dimup] = in[ju + sc * dimup];
  }
}

// Start vector of a spin-DOWN channel into c->d_lx: c / c^+ on the dw word maps every target column onto ONE column
// of the state with a sign (the up word is a spectator), so it is a signed column permutation.  On one rank the
// columns are read in place; on a sharded state each rank receives the columns it needs from their owners in one
// grouped exchange (the reference gathers the whole state on the master and scatters the result,
// ED_GF_NORMAL.f90:184-216, ED_EIGENSPACE.f90:540-549).
static int gf_start_dw(edgpu_ctx *c, int iorb, int add) {
  const int P = c->nranks, me = c->rank;
  const uint32_t bit = 1u << (iorb - 1);
  const int64_t sdimdw = c->h_binom[c->ns * EDGPU_BINOM_LD + c->gs_ndw];
  const int64_t dimup = c->dimup;                                  // the up basis is the state's
  std::vector<int32_t> tmap((size_t)c->dimdw);
  CK(cudaMemcpy(tmap.data(), c->dw.d_map, tmap.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
  std::vector<int64_t> gq((size_t)P), goff((size_t)P), tq((size_t)P), toff((size_t)P);
  for (int p = 0; p < P; p++) { edgpu_split(sdimdw, P, p, &gq[p], &goff[p]); edgpu_split(c->dimdw, P, p, &tq[p], &toff[p]); }
  auto gs_owner = [&](int64_t sc) { int p = 0; while (p + 1 < P && sc >= goff[(size_t)p + 1]) p++; return p; };
  // source column (in the state's dw basis) and sign of target column t, -1 if the operator annihilates it
  auto source_of = [&](int64_t t, double &sg) -> int64_t {
    const uint32_t r = (uint32_t)tmap[(size_t)t];
    const bool has = (r & bit) != 0;
    if (add ? !has : has) return -1;
    const uint32_t m = add ? (r & ~bit) : (r | bit);
    sg = hd_sign_below(m, iorb);
    return hd_rank(m, c->h_binom);
  };
  // what I receive, target column by target column (segments ordered by owner), and what every peer receives from me
  std::vector<int> src((size_t)c->qdw, -1);
  std::vector<double> sgn((size_t)c->qdw, 0.0);
  std::vector<int64_t> rcnt((size_t)P, 0), roff((size_t)P, 0), scnt((size_t)P, 0), soff((size_t)P, 0);
  std::vector<int> sendcols;
  if (P == 1) {
    for (int64_t jl = 0; jl < c->qdw; jl++) { double sg = 0.0; const int64_t sc = source_of(jl, sg); src[(size_t)jl] = (int)sc; sgn[(size_t)jl] = sg; }
  } else {
    std::vector<int> own((size_t)c->qdw, -1);
    for (int64_t jl = 0; jl < c->qdw; jl++) {
      double sg = 0.0;
      const int64_t sc = source_of(toff[(size_t)me] + jl, sg);
      if (sc < 0) continue;
      own[(size_t)jl] = gs_owner(sc); sgn[(size_t)jl] = sg;
      rcnt[(size_t)own[(size_t)jl]]++;
    }
    for (int p = 1; p < P; p++) roff[(size_t)p] = roff[(size_t)p - 1] + rcnt[(size_t)p - 1];
    std::vector<int64_t> fill(roff);
    for (int64_t jl = 0; jl < c->qdw; jl++) if (own[(size_t)jl] >= 0) src[(size_t)jl] = (int)fill[(size_t)own[(size_t)jl]]++;
    for (int q = 0; q < P; q++) {                                  // rank q's requests to me, in q's target order
      soff[(size_t)q] = (int64_t)sendcols.size();
      for (int64_t jl = 0; jl < tq[(size_t)q]; jl++) {
        double sg;
        const int64_t sc = source_of(toff[(size_t)q] + jl, sg);
        if (sc >= 0 && gs_owner(sc) == me) sendcols.push_back((int)(sc - goff[(size_t)me]));
      }
      scnt[(size_t)q] = (int64_t)sendcols.size() - soff[(size_t)q];
    }
  }
  int *d_src = nullptr, *d_idx = nullptr;
  double *d_sgn = nullptr, *d_sbuf = nullptr, *d_rbuf = nullptr;
  int rc = EDGPU_OK;
  auto fail = [&](const char *what) { rc = edgpu_set_err(EDGPU_ERR_CUDA, "gf_start_dw: %s", what); };
  const size_t nq = (size_t)std::max<int64_t>(c->qdw, 1);
  if (cudaMalloc(&d_src, nq * sizeof(int)) != cudaSuccess || cudaMalloc(&d_sgn, nq * sizeof(double)) != cudaSuccess) fail("cudaMalloc");
  if (!rc && (cudaMemcpy(d_src, src.data(), (size_t)c->qdw * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ||
              cudaMemcpy(d_sgn, sgn.data(), (size_t)c->qdw * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess)) fail("upload");
  const double *S = c->d_gs;
  if (!rc && P > 1) {
    const int64_t nsend = (int64_t)sendcols.size(), nrecv = roff[(size_t)P - 1] + rcnt[(size_t)P - 1];
    if (cudaMalloc(&d_idx, (size_t)std::max<int64_t>(nsend, 1) * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&d_sbuf, (size_t)std::max<int64_t>(nsend * dimup, 1) * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&d_rbuf, (size_t)std::max<int64_t>(nrecv * dimup, 1) * sizeof(double)) != cudaSuccess) fail("cudaMalloc (exchange)");
    if (!rc && nsend) {
      if (cudaMemcpy(d_idx, sendcols.data(), (size_t)nsend * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) fail("upload (exchange)");
      dim3 grid((unsigned)((dimup + 255) / 256), (unsigned)std::min<int64_t>(nsend, 32768));
      if (!rc) { k_gather_cols<<<grid, 256, 0, c->stream>>>(c->d_gs, d_idx, dimup, nsend, d_sbuf); c->launches++; }
    }
    if (!rc) {
      for (int p = 0; p < P; p++) { soff[(size_t)p] *= dimup; scnt[(size_t)p] *= dimup; roff[(size_t)p] *= dimup; rcnt[(size_t)p] *= dimup; }
      rc = comm_exchange(c, d_sbuf, soff.data(), scnt.data(), d_rbuf, roff.data(), rcnt.data());
    }
    S = d_rbuf;
  }
  if (!rc && c->qdw > 0) {
    dim3 grid((unsigned)((dimup + 255) / 256), (unsigned)std::min<int64_t>(c->qdw, 32768));
    k_gf_start_dw<<<grid, 256, 0, c->stream>>>(S, d_src, d_sgn, dimup, c->qdw, c->d_lx);
    c->launches++;
    if (cudaGetLastError() != cudaSuccess) fail("k_gf_start_dw launch");
  }
  if (cudaStreamSynchronize(c->stream) != cudaSuccess && !rc) fail("sync");
  cudaFree(d_src); cudaFree(d_sgn); cudaFree(d_idx); cudaFree(d_sbuf); cudaFree(d_rbuf);
  return rc;
}

// One chain of a batch: its own Lanczos vectors, scalars, coefficient arrays and stream.  The recurrences of all the
// channels that share a target sector advance in lock step, one stream each, so kernels of different chains overlap on
// the device while the sector's factors, group records and packed entries are built and read once.
struct ChainWork {
  cudaStream_t stream = nullptr;
  double *lx = nullptr, *lp = nullptr, *lt = nullptr, *partials = nullptr, *alanc = nullptr, *blanc = nullptr;
  LancState *st = nullptr;
};
struct ChainBind {                                                 // the context's single-chain state, swapped while a chain is bound
  cudaStream_t stream; double *lx, *lp, *lt, *partials, *alanc, *blanc; LancState *st; int cap;
};
static void chain_bind(edgpu_ctx *c, ChainWork &w, ChainBind &keep, int cap) {
  keep = {c->stream, c->d_lx, c->d_lp, c->d_lt, c->d_partials, c->d_alanc, c->d_blanc, c->d_st, c->lanc_cap};
  c->stream = w.stream; c->d_lx = w.lx; c->d_lp = w.lp; c->d_lt = w.lt; c->d_partials = w.partials;
  c->d_alanc = w.alanc; c->d_blanc = w.blanc; c->d_st = w.st; c->lanc_cap = cap;
}
static void chain_unbind(edgpu_ctx *c, ChainWork &w, const ChainBind &keep) {
  w.lx = c->d_lx; w.lp = c->d_lp;                                   // the step rotates the two
  c->stream = keep.stream; c->d_lx = keep.lx; c->d_lp = keep.lp; c->d_lt = keep.lt; c->d_partials = keep.partials;
  c->d_alanc = keep.alanc; c->d_blanc = keep.blanc; c->d_st = keep.st; c->lanc_cap = keep.cap;
}
static void chain_free(ChainWork &w) {
  cudaFree(w.lx); cudaFree(w.lp); cudaFree(w.lt); cudaFree(w.partials); cudaFree(w.alanc); cudaFree(w.blanc); cudaFree(w.st);
  if (w.stream) cudaStreamDestroy(w.stream);
  w = ChainWork();
}
static int chain_alloc(edgpu_ctx *c, ChainWork &w, int nl) {
  const size_t nb = ((size_t)std::max<int64_t>(c->nloc, 1) + 2) * sizeof(double);
  CK(cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
  CK(cudaMalloc(&w.lx, nb)); CK(cudaMalloc(&w.lp, nb)); CK(cudaMalloc(&w.lt, nb));
  CK(cudaMalloc(&w.partials, 4096 * sizeof(double)));
  CK(cudaMalloc(&w.alanc, (size_t)(nl + 2) * sizeof(double))); CK(cudaMalloc(&w.blanc, (size_t)(nl + 2) * sizeof(double)));
  CK(cudaMalloc(&w.st, sizeof(LancState)));
  CK(cudaMemsetAsync(w.alanc, 0, (size_t)(nl + 2) * sizeof(double), w.stream));
  CK(cudaMemsetAsync(w.blanc, 0, (size_t)(nl + 2) * sizeof(double), w.stream));
  CK(cudaMemsetAsync(w.lp, 0, nb, w.stream));
  CK(cudaMemsetAsync(w.st, 0, sizeof(LancState), w.stream));
  return EDGPU_OK;
}

// start vector of one channel into c->d_lx (on c->stream)
static int gf_start_vector(edgpu_ctx *c, int iorb, int ispin, int add) {
  if (ispin == 2) return gf_start_dw(c, iorb, add);
  const int64_t sdimup = c->h_binom[c->ns * EDGPU_BINOM_LD + c->gs_nup];
  dim3 grid((unsigned)((c->dimup + 255) / 256), (unsigned)(c->qdw < 32768 ? c->qdw : 32768));
  k_gf_start<<<grid, 256, 0, c->stream>>>(c->up.d_map, c->dw.d_map, c->dimup, c->qdw, c->coloff, sdimup,
                                          c->d_gs, iorb, add, c->d_binom, c->d_lx);
  CKL(c);
  return EDGPU_OK;
}

// reference bookkeeping of sp_lanc_tridiag on raw device coefficients: alanc(k)=a_k ; if |b_k|<threshold exit ; blanc(k+1)=b_k
static void tridiag_bookkeeping(int nlanc, double threshold, const double *a, const double *b, double *alanc, double *blanc) {
  for (int k = 0; k < nlanc; k++) { alanc[k] = 0.0; blanc[k] = 0.0; }
  for (int k = 0; k < nlanc; k++) {
    alanc[k] = a[k];
    const double bk = b[k + 1];
    if (!(fabs(bk) >= threshold)) break;          // also stops on NaN
    if (k + 1 < nlanc) blanc[k + 1] = bk;
  }
}

// all the chains of one target sector (live in c), batched: lock-step recurrences on one stream per chain
// start(ch): the start vector of channel ch into c->d_lx on c->stream
static int gf_chains_batched(edgpu_ctx *c, const std::vector<int> &grp, const std::function<int(int)> &start,
                             int nl, int nlanc_max, double threshold, double *norm2, int *nlanc, double *alanc, double *blanc) {
  const int nb = (int)grp.size();
  std::vector<ChainWork> work((size_t)nb);
  std::vector<double> ha((size_t)nb * (nl + 2)), hb((size_t)nb * (nl + 2)), hn((size_t)nb);
  ChainBind keep;
  int rc = EDGPU_OK;
  CK(cudaStreamSynchronize(c->stream));                            // the sector build is complete before the chains' streams start
  for (int q = 0; q < nb && !rc; q++) {
    rc = chain_alloc(c, work[(size_t)q], nl);
    if (rc) break;
    chain_bind(c, work[(size_t)q], keep, nl + 2);
    rc = start(grp[(size_t)q]);
    if (!rc) rc = lanczos_norm_start(c);
    chain_unbind(c, work[(size_t)q], keep);
  }
  for (int k = 0; k < nl && !rc; k++)
    for (int q = 0; q < nb && !rc; q++) {
      chain_bind(c, work[(size_t)q], keep, nl + 2);
      rc = lanczos_step(c, k);
      chain_unbind(c, work[(size_t)q], keep);
    }
  for (int q = 0; q < nb && !rc; q++) {
    ChainWork &w = work[(size_t)q];
    if (cudaMemcpyAsync(&ha[(size_t)q * (nl + 2)], w.alanc, (size_t)(nl + 1) * sizeof(double), cudaMemcpyDeviceToHost, w.stream) != cudaSuccess ||
        cudaMemcpyAsync(&hb[(size_t)q * (nl + 2)], w.blanc, (size_t)(nl + 2) * sizeof(double), cudaMemcpyDeviceToHost, w.stream) != cudaSuccess ||
        cudaMemcpyAsync(&hn[(size_t)q], &w.st->norm2, sizeof(double), cudaMemcpyDeviceToHost, w.stream) != cudaSuccess)
      rc = edgpu_set_err(EDGPU_ERR_CUDA, "gf_chains: coefficient read-back failed");
  }
  for (int q = 0; q < nb; q++)
    if (work[(size_t)q].stream && cudaStreamSynchronize(work[(size_t)q].stream) != cudaSuccess && !rc)
      rc = edgpu_set_err(EDGPU_ERR_CUDA, "gf_chains: %s", cudaGetErrorString(cudaGetLastError()));
  for (int q = 0; q < nb && !rc; q++) {
    const int ch = grp[(size_t)q];
    tridiag_bookkeeping(nl, threshold, &ha[(size_t)q * (nl + 2)], &hb[(size_t)q * (nl + 2)], alanc + (size_t)ch * nlanc_max, blanc + (size_t)ch * nlanc_max);
    norm2[ch] = hn[(size_t)q];
    nlanc[ch] = nl;
  }
  for (int q = 0; q < nb; q++) chain_free(work[(size_t)q]);
  return rc;
}

extern "C" int edgpu_gf_chains(edgpu_ctx *c, int nchains, const int *iorb, const int *ispin,
                               const int *addrem, int nlanc_max, double threshold,
                               double *norm2, int *nlanc, double *alanc, double *blanc) {
  if (!c || !c->d_gs) return edgpu_set_err(EDGPU_ERR_INVALID, "gf_chains: no state set (edgpu_gf_set_state)");
  if (!c->hp.ed_total_ud) return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "ed_total_ud = F: GF chains of an orbital-resolved state are not built");
  if (c->hstatus) return edgpu_set_err(EDGPU_ERR_INVALID, "gf_chains: a sector is live; call delete_Hv_sector first");
  CK(cudaSetDevice(c->device));
  std::vector<int> done((size_t)nchains, 0);
  for (int ch = 0; ch < nchains; ch++) {
    norm2[ch] = 0.0; nlanc[ch] = 0;
    for (int k = 0; k < nlanc_max; k++) { alanc[(size_t)ch * nlanc_max + k] = 0.0; blanc[(size_t)ch * nlanc_max + k] = 0.0; }
    if (iorb[ch] < 1 || iorb[ch] > c->dp.norb || ispin[ch] < 1 || ispin[ch] > 2 || (addrem[ch] != 1 && addrem[ch] != -1))
      return edgpu_set_err(EDGPU_ERR_INVALID, "gf_chains: bad channel %d", ch);
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
    std::vector<int> grp;                                   // every channel sharing this target sector
    for (int ch2 = ch; ch2 < nchains; ch2++) {
      if (done[ch2]) continue;
      int n2 = c->gs_nup + (ispin[ch2] == 1 ? addrem[ch2] : 0), d2 = c->gs_ndw + (ispin[ch2] == 2 ? addrem[ch2] : 0);
      if (n2 != jnup || d2 != jndw) continue;
      done[ch2] = 1;
      grp.push_back(ch2);
    }
    const int64_t jdim = c->dimup * c->dimdw;
    const int nl = (int)std::min<int64_t>(jdim, nlanc_max);   // nlanc=min(jdim,lanc_nGFiter), :219
    int rc = EDGPU_OK;
    // the batch needs 3 vectors per chain; sharded runs share one halo buffer per context, so their chains go one by one
    const bool batch = c->nranks == 1 && grp.size() > 1 && !c->opt_no_batch &&
                       (double)grp.size() * 3.0 * 8.0 * (double)c->nloc < 60e9;
    if (batch) {
      rc = gf_chains_batched(c, grp, [&](int ch2) { return gf_start_vector(c, iorb[ch2], ispin[ch2], addrem[ch2] == 1 ? 1 : 0); },
                             nl, nlanc_max, threshold, norm2, nlanc, alanc, blanc);
    } else {
      for (size_t q = 0; q < grp.size() && !rc; q++) {
        const int ch2 = grp[q];
        rc = lanczos_begin(c, nl);
        if (!rc) rc = gf_start_vector(c, iorb[ch2], ispin[ch2], addrem[ch2] == 1 ? 1 : 0);
        if (!rc) rc = tridiag_device(c, nl, threshold, alanc + (size_t)ch2 * nlanc_max, blanc + (size_t)ch2 * nlanc_max);
        if (!rc && cudaMemcpy(&norm2[ch2], &c->d_st->norm2, sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess)
          rc = edgpu_set_err(EDGPU_ERR_CUDA, "norm2 read-back failed");
        if (!rc) nlanc[ch2] = nl;
      }
    }
    edgpu_delete_hv_sector(c);
    if (rc) return rc;
  }
  return EDGPU_OK;
}

// ---- susceptibility chains (ED_GF_CHISPIN.f90:114-415, ED_GF_CHIDENS.f90:111-426) ---------------------------------------
// vvinit = O|gs> with a DIAGONAL operator of the impurity occupations, in the state's own sector: spin
// O = 1/2 sum_a (n_up,a - n_dw,a), density O = sum_a (n_up,a + n_dw,a), a over the orbitals of `mask` (one orbital: main;
// two: mix, S_i + S_j; all: tot).  No gathered vector, no master-only loop: every rank scales its own shard.
__global__ void k_chi_start(const int32_t *__restrict__ map_up, const int32_t *__restrict__ map_dw, int64_t dimup, int64_t qdw,
                            int64_t coloff, const double *__restrict__ gs, uint32_t mask, int kind, double *__restrict__ out) {
  for (int64_t jl = blockIdx.y; jl < qdw; jl += gridDim.y) {
    const int ndw = __popc((uint32_t)map_dw[coloff + jl] & mask);
    for (int64_t ju = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; ju < dimup; ju += (int64_t)gridDim.x * blockDim.x) {
      const int nup = __popc((uint32_t)map_up[ju] & mask);
      const double w = kind == 0 ? 0.5 * (double)(nup - ndw) : (double)(nup + ndw);
      out[ju + jl * dimup] = w * gs[ju + jl * dimup];
    }
  }
}

extern "C" int edgpu_chi_chains(edgpu_ctx *c, int kind, int nchains, const int *iorb, const int *jorb, int nlanc_max,
                                double threshold, double *norm2, int *nlanc, double *alanc, double *blanc) {
  if (!c || !c->d_gs) return edgpu_set_err(EDGPU_ERR_INVALID, "chi_chains: no state set (edgpu_gf_set_state)");
  if (!c->hp.ed_total_ud) return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "ed_total_ud = F: susceptibility chains of an orbital-resolved state are not built");
  if (c->hstatus) return edgpu_set_err(EDGPU_ERR_INVALID, "chi_chains: a sector is live; call delete_Hv_sector first");
  if (kind != 0 && kind != 1) return edgpu_set_err(EDGPU_ERR_INVALID, "chi_chains: kind is 0 (spin) or 1 (density)");
  CK(cudaSetDevice(c->device));
  std::vector<uint32_t> mask((size_t)nchains);
  for (int ch = 0; ch < nchains; ch++) {
    norm2[ch] = 0.0; nlanc[ch] = 0;
    for (int k = 0; k < nlanc_max; k++) { alanc[(size_t)ch * nlanc_max + k] = 0.0; blanc[(size_t)ch * nlanc_max + k] = 0.0; }
    const int io = iorb[ch], jo = jorb[ch];
    if (io == 0) mask[(size_t)ch] = (1u << c->dp.norb) - 1u;                                  // _tot_main
    else if (io >= 1 && io <= c->dp.norb && jo >= 1 && jo <= c->dp.norb) mask[(size_t)ch] = (1u << (io - 1)) | (1u << (jo - 1));   // _main / _mix_main
    else return edgpu_set_err(EDGPU_ERR_INVALID, "chi_chains: bad channel %d", ch);
  }
  int isector;
  TRY(edgpu_get_sector(c, c->gs_nup, c->gs_ndw, &isector));
  TRY(edgpu_build_hv_sector(c, isector));
  int rc = EDGPU_OK;
  if (c->nloc != c->gs_nloc) rc = edgpu_set_err(EDGPU_ERR_INVALID, "chi_chains: the state was set for another rank layout");
  const int64_t dim = c->dimup * c->dimdw;
  const int nl = (int)std::min<int64_t>(dim, nlanc_max);   // nlanc=min(idim,lanc_nGFiter), ED_GF_CHISPIN.f90:176
  auto start = [&](int ch) -> int {
    dim3 grid((unsigned)((c->dimup + 255) / 256), (unsigned)(c->qdw < 32768 ? std::max<int64_t>(c->qdw, 1) : 32768));
    k_chi_start<<<grid, 256, 0, c->stream>>>(c->up.d_map, c->dw.d_map, c->dimup, c->qdw, c->coloff, c->d_gs, mask[(size_t)ch], kind, c->d_lx);
    CKL(c);
    return EDGPU_OK;
  };
  std::vector<int> grp((size_t)nchains);
  for (int ch = 0; ch < nchains; ch++) grp[(size_t)ch] = ch;
  const bool batch = !rc && c->nranks == 1 && nchains > 1 && !c->opt_no_batch && (double)nchains * 3.0 * 8.0 * (double)c->nloc < 60e9;
  if (!rc && batch) {
    rc = gf_chains_batched(c, grp, start, nl, nlanc_max, threshold, norm2, nlanc, alanc, blanc);
  } else {
    for (int ch = 0; ch < nchains && !rc; ch++) {
      rc = lanczos_begin(c, nl);
      if (!rc) rc = start(ch);
      if (!rc) rc = tridiag_device(c, nl, threshold, alanc + (size_t)ch * nlanc_max, blanc + (size_t)ch * nlanc_max);
      if (!rc && cudaMemcpy(&norm2[ch], &c->d_st->norm2, sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess)
        rc = edgpu_set_err(EDGPU_ERR_CUDA, "norm2 read-back failed");
      if (!rc) nlanc[ch] = nl;
    }
  }
  edgpu_delete_hv_sector(c);
  return rc;
}

// add_to_lanczos_spinChi (ED_GF_CHISPIN.f90:434-488) == add_to_lanczos_densChi (ED_GF_CHIDENS.f90:436-489), T = 0:
// chi_iv[0..lmats] on the bosonic vm, chi_tau[0..ltau], chi_w[lreal] (interleaved complex); accumulates.
extern "C" int edgpu_add_to_lanczos_chi(double norm2, double zeta, double ei, double beta, const double *alanc, const double *blanc,
                                        int nlanc, const double *vm, int lmats, double *chi_iv, const double *tau, int ltau,
                                        double *chi_tau, const double *vr, int lreal, double eps, double *chi_w) {
  if (nlanc < 1) return edgpu_set_err(EDGPU_ERR_INVALID, "add_to_lanczos_chi: nlanc < 1");
  std::vector<double> diag, z;
  TRY(tridiag_eig(nlanc, alanc, blanc, diag, z));
  std::complex<double> *cw = reinterpret_cast<std::complex<double> *>(chi_w);
  const double pesof = norm2 / zeta;                              // pesoBZ = 1 at T = 0
  for (int j = 0; j < nlanc; j++) {
    const double de = diag[j] - ei;
    const double z1 = z[0 + (size_t)nlanc * j];
    const double peso = pesof * (z1 * z1);
    const double bose = 1.0 - exp(-beta * de);
    if (chi_iv) {
      if (beta * de > 1e-3) chi_iv[0] += peso * 2 * bose / de;
      for (int i = 1; i <= lmats; i++) chi_iv[i] += peso * bose * 2.0 * de / (vm[i] * vm[i] + de * de);
    }
    if (chi_tau) for (int i = 0; i <= ltau; i++) chi_tau[i] += exp(-tau[i] * de) * peso;
    if (chi_w) for (int i = 0; i < lreal; i++) {
      const std::complex<double> w(vr[i], eps);
      cw[i] -= peso * bose * (1.0 / (w - de) - 1.0 / (w + de));
    }
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
