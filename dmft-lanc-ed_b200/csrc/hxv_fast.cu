// hxv_fast.cu -- the engine's fast H*v path: two HBM passes, each staged through shared memory by the TMA
// engine (cp.async.bulk / cp.async.bulk.tensor + mbarrier, double buffered, persistent CTAs, one CTA per SM,
// a dedicated producer warp and full/empty barriers instead of CTA-wide barriers).
//
//   y = Hd o x + Hup x + x Hdw^T       x(i_up, i_dw) column-major, i_up contiguous
//
//   pass 1  k_srow : y = Hd o x + x Hdw^T (write only).  Tile = 32 consecutive i_up rows x one chunk of i_dw
//                    columns (2-D tensor-map TMA boxes); lanes run along i_up so every shared-memory access is
//                    unit stride, and the dw hops are generated from the bit structure of the star geometry
//                    (Norb = 1: every hop is impurity bit 0 <-> bath bit k): the columns of one "low group"
//                    (same high bits, LR low bits) live in registers, hops among the low bits are register to
//                    register, a hop on a high bit moves the whole group to ONE other group whose base column
//                    comes from a Lin table.  No per-element index data at all.  Sources outside the tile come
//                    from L2 -- or, when the vector is sharded over GPUs, from the owner rank's copy over NVLink
//                    (peer-mapped symmetric slab, comm.cu); the few hops that touch a low group cut by a rank
//                    boundary are left to k_sfix.
//   pass 2  k_fcol : y += F x(:, j) with one WHOLE column per stage in shared memory (a contiguous 8*DimUp byte
//                    bulk copy).  F is any one-spin factor (spH0ups / spH0dws, ED_HAMILTONIAN/stored/H_up.f90,
//                    H_dw.f90) in a packed ELL form streamed from L2; every gather hits shared memory.  MODE 2
//                    fuses the first Lanczos vector update (w = s*Hx - c*x_prev, alpha partials) into the epilogue.
//           k_fcol2: the same for columns that exceed one SM (Ns = 18): a 2-CTA cluster holds the column, the
//                    other half is read through distributed shared memory.
//
// The factor values are exactly the reference's V_k * sg1 * sg2 (stored/H_up.f90:55-81); only the
// order of the floating-point sums differs.
#include <cuda.h>

#include <algorithm>
#include <map>
#include <utility>
#include <vector>

#include "engine.h"

#define F_COL_BITS 20
#define F_COL_MASK 0xFFFFFu
#define F_VID_MASK 0x7FFu
#define F_MAXVALS 256
#define FCOL_THREADS 1024
// consumer threads per CTA (+1 producer warp): LR=5 -> 512 threads x 128 registers, LR=4 -> 768 x 85
__host__ __device__ constexpr int srow_consumers(int LR) { return LR >= 5 ? 480 : 736; }
#define SROW_R 32
#ifndef SROW_NB_FAR
#define SROW_NB_FAR 3             // high-bit hops whose loads are fused (far sources: L2 / peer GPU)
#endif
#ifndef SROW_NB_IN
#define SROW_NB_IN 1              // ... for sources inside the shared-memory tile
#endif
#define SROW_BC 32                // columns per TMA box (32 rows x 32 columns x 8 B = 8 KB per copy)
#define SMEM_LIMIT 232448        // 227 KB per CTA on sm_100
#define JHI_COLMASK 0xFFFFF      // Lin table entry: first column | owner rank << 20 | (cut by a rank boundary) << 30
#define JHI_CUT 0x40000000
#define SROW_MAXP 8              // peer reads: ranks of one NVLink domain

struct FastFactor {
  int W = 0, WT = 0, nvals = 0;
  int64_t n = 0;
  uint32_t *d_ell = nullptr;     // [WT][n] slot-major, padding entries point at column n (zero slot)
  double *d_vtab = nullptr;      // [nvals] magnitudes, vtab[0] = 0
  uint16_t *d_ell16 = nullptr;   // uniform factors with n < 32768: column | sign << 15 (halves the L2 index traffic);
                                 // [n][8] row-major when WT == 8 (one 16-byte load per row), else [WT][n]
  bool uniform = false;          // every stored magnitude identical -> y = v * sum(+-x)
  double vuni = 0.0;
};

struct SRowPlan {
  bool ok = false;
  int LR = 0, nhigh = 0, ngroups = 0, nchunks = 0, cmax = 0;
  int32_t *d_jhi = nullptr;      // [2^nhigh] first column of group h, -1 if the group is empty
  uint16_t *d_grp = nullptr;     // [ngroups] high words of the non-empty groups, ascending
  int4 *d_chunks = nullptr;      // [nchunks] (group begin, group end, column begin, column end)
  double *d_dr0 = nullptr, *d_dr1 = nullptr;   // direct mode: dfac_up[row] (+ Uloc when the up impurity is occupied)
  // sharded vector: hops that touch a low group cut by a rank boundary are left to a small fix-up kernel
  int nfix = 0;                  // target columns of the fix-up
  int *d_ftptr = nullptr, *d_ftcol = nullptr, *d_feown = nullptr, *d_fesrc = nullptr;
  unsigned char *d_ftinit = nullptr;
  double *d_feval = nullptr;
  int coloffs[65];
  double vk[EDGPU_MAX_SITES];    // V_k of the dw spin, k = bath bit (1-based site k+1)
  size_t smem = 0;
};

struct FastPlan {
  FastFactor ff[2];
  bool col_ok[2] = {false, false};
  size_t col_smem[2] = {0, 0};
  bool col2_ok[2] = {false, false};   // two-CTA cluster variant (half a column per SM)
  size_t col2_smem[2] = {0, 0};
  SRowPlan sr;
};

// ---------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D bulk copy (TMA engine; SASS: UBLKCP / SYNCS)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tm, int c0, int c1, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ double flip_sign(double v, uint32_t signbit31) {
  return __hiloint2double(__double2hiint(v) ^ (int)signbit31, __double2loint(v));
}

// ---------------------------------------------------------------------------------------------
// pass 1: whole-column kernel
// ---------------------------------------------------------------------------------------------
struct FColArgs {
  const double *x;
  double *y;
  int n;                         // factor dimension (even)
  int64_t ncols, coloff;
  const uint32_t *ell;
  const uint16_t *ell16;
  const double *vtab;
  int nvals;
  double vuni;
  const double *diag;            // DIAG == 1
  const double *dfac_c, *dfac_s; // DIAG == 2
  const int32_t *map_c, *map_s;
  int norb;
  double uloc[EDGPU_MAX_ORB];
  double ust;
  // MODE == 2 (Lanczos epilogue): w = sx*(y + F x) - cprev*xp, written over xp; partials of (sx*x).w
  double *xp;
  const LancState *st;
  double *partials;
};

// MODE: 0 = y = F x, 1 = y += F x, 2 = y += F x fused with the first Lanczos vector update (y is only read)
// UNI: 0 = general (value table), 1 = uniform magnitude with 4-byte entries, 2 = uniform with 2-byte entries
template <int WT, int DIAG, int UNI, int MODE>
__global__ void __launch_bounds__(FCOL_THREADS, 1) k_fcol(FColArgs a) {
  constexpr bool ACC = MODE >= 1;
  constexpr int NCW = FCOL_THREADS / 32 - 1;                      // 31 consumer warps + 1 producer warp
  constexpr int NCT = NCW * 32;
  extern __shared__ __align__(128) unsigned char smraw[];
  const int n = a.n, ld = n + 2;
  double *buf0 = reinterpret_cast<double *>(smraw);
  double *buf1 = buf0 + ld;
  double *vtab = buf1 + ld;
  uint64_t *bar = reinterpret_cast<uint64_t *>(vtab + F_MAXVALS);  // full[2], empty[2]
  double *red = reinterpret_cast<double *>(bar + 4);             // [32] block reduction (MODE == 2)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  double lsx = 0.0, lcp = 0.0, lsum = 0.0;
  if (MODE == 2) { lsx = a.st->sx; lcp = a.st->cprev; }
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_init(&bar[2], NCW);
    mbar_init(&bar[3], NCW);
    fence_barrier_init();
    buf0[n] = 0.0; buf0[n + 1] = 0.0;      // zero slot for the ELL padding entries
    buf1[n] = 0.0; buf1[n + 1] = 0.0;
  }
  if (UNI == 0)
    for (int i = tid; i < a.nvals; i += FCOL_THREADS) vtab[i] = a.vtab[i];
  __syncthreads();
  const uint32_t colbytes = (uint32_t)n * 8u;
  const int64_t G = gridDim.x;
  if (warp == NCW) {
    // ---- producer warp: one whole column per stage; a stage is refilled as soon as every consumer warp
    // has released it (no CTA-wide barrier anywhere in the loop)
    if (lane == 0) {
      for (int64_t it = 0;; it++) {
        const int64_t j = blockIdx.x + it * G;
        if (j >= a.ncols) break;
        const int b = (int)(it & 1);
        if (it >= 2) mbar_wait(&bar[2 + b], (uint32_t)(((it >> 1) - 1) & 1));
        fence_proxy_async();
        mbar_expect_tx(&bar[b], colbytes);
        const char *src = reinterpret_cast<const char *>(a.x + j * (int64_t)n);
        char *dst = reinterpret_cast<char *>(b ? buf1 : buf0);
        for (uint32_t off = 0; off < colbytes; off += 32768u)
          bulk_g2s(dst + off, src + off, min(32768u, colbytes - off), &bar[b]);
      }
    }
  } else {
    for (int64_t it = 0;; it++) {
      const int64_t j = blockIdx.x + it * G;
      if (j >= a.ncols) break;
      const int b = (int)(it & 1);
      const double *xs = b ? buf1 : buf0;
      double dcol = 0.0;
      uint32_t ms = 0;
      if (DIAG == 2) { dcol = a.dfac_s[a.coloff + j]; ms = (uint32_t)a.map_s[a.coloff + j]; }
      double *yc = a.y + j * (int64_t)n;
      uint32_t en[WT];
      auto load_ell = [&](int row) {
        if (UNI == 2 && WT == 8) {                                 // 8 two-byte entries = one LDG.128
          const uint4 q = __ldg(reinterpret_cast<const uint4 *>(a.ell16) + row);
          en[0] = q.x & 0xFFFFu; en[1] = q.x >> 16; en[2] = q.y & 0xFFFFu; en[3] = q.y >> 16;
          en[4 % WT] = q.z & 0xFFFFu; en[5 % WT] = q.z >> 16; en[6 % WT] = q.w & 0xFFFFu; en[7 % WT] = q.w >> 16;
        } else {
#pragma unroll
          for (int s = 0; s < WT; s++) en[s] = (UNI == 2) ? (uint32_t)__ldg(a.ell16 + (size_t)s * n + row) : __ldg(a.ell + (size_t)s * n + row);
        }
      };
      if (tid < n) load_ell(tid);                                  // independent of the column: issued before the wait
      double ynext = 0.0;
      if (ACC && tid < n) ynext = __ldcs(yc + tid);
      mbar_wait(&bar[b], (uint32_t)((it >> 1) & 1));
      for (int r = tid; r < n; r += NCT) {
        uint32_t e[WT];
#pragma unroll
        for (int s = 0; s < WT; s++) e[s] = en[s];
        const double yold = ynext;
        if (r + NCT < n) {                                         // software prefetch of the next row's inputs
          load_ell(r + NCT);
          if (ACC) ynext = __ldcs(yc + r + NCT);
        }
        double acc0 = 0.0;
        if (DIAG == 1) acc0 = __ldcs(a.diag + j * (int64_t)n + r) * xs[r];
        if (DIAG == 2) {
          double d = __ldg(a.dfac_c + r) + dcol;
          const uint32_t mc = (uint32_t)__ldg(a.map_c + r);
          for (int o = 0; o < a.norb; o++)
            if ((mc >> o) & 1u)
              for (int q = 0; q < a.norb; q++)
                if ((ms >> q) & 1u) d += (o == q) ? a.uloc[o] : a.ust;
          acc0 = d * xs[r];
        }
        double acc = 0.0;
#pragma unroll
        for (int s = 0; s < WT; s++) {
          if (UNI == 2) { acc += flip_sign(xs[e[s] & 0x7FFFu], (e[s] & 0x8000u) << 16); continue; }
          const double xv = xs[e[s] & F_COL_MASK];
          if (UNI) acc += flip_sign(xv, e[s] & 0x80000000u);
          else acc = fma(flip_sign(vtab[(e[s] >> F_COL_BITS) & F_VID_MASK], e[s] & 0x80000000u), xv, acc);
        }
        if (UNI) acc0 = fma(a.vuni, acc, acc0); else acc0 += acc;
        if (ACC) acc0 += yold;
        if (MODE == 2) {
          double *wp = a.xp + j * (int64_t)n + r;
          const double w = lsx * acc0 - lcp * __ldcs(wp);
          __stcs(wp, w);
          lsum = fma(lsx * xs[r], w, lsum);
        } else {
          __stcs(yc + r, acc0);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[2 + b]);                     // this warp is done with the stage
    }
  }
  if (MODE == 2) {                                               // deterministic: fixed order inside the CTA, one partial per CTA
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    if (lane == 0) red[warp] = lsum;
    __syncthreads();
    if (tid < 32) {
      double r2 = red[tid];
      for (int o = 16; o > 0; o >>= 1) r2 += __shfl_xor_sync(0xffffffffu, r2, o);
      if (tid == 0) a.partials[blockIdx.x] = r2;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// pass 1 for columns that do not fit one SM's shared memory (Ns = 18: 389 KB): a CLUSTER of two CTAs holds
// the column, one half each; a source in the other half is read through distributed shared memory
// (mapa + ld.shared::cluster).  In the sorted basis the halves are (nearly) the two values of the top bit, so
// only the hops on that bit cross (1/17 of the gathers at Ns = 18).  Single stage per CTA: the half fills the SM.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local), "r"(rank));
  return ra;
}
__device__ __forceinline__ double ld_dsmem(uint32_t addr) {
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}

template <int WT, int UNI, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FCOL_THREADS, 1) k_fcol2(FColArgs a) {
  extern __shared__ __align__(128) unsigned char smraw[];
  constexpr bool ACC = MODE >= 1;
  const int n = a.n;
  const int nh0 = ((n >> 1) + 1) & ~1;                            // rows [0, nh0) on CTA 0, [nh0, n) on CTA 1
  const uint32_t me = cluster_ctarank();
  const int r0 = me ? nh0 : 0, nr = me ? n - nh0 : nh0;
  const int r0p = me ? 0 : nh0, nrp = me ? nh0 : n - nh0;         // the partner's rows
  const int hmax = nh0 + 2;
  double *buf = reinterpret_cast<double *>(smraw);                // [hmax]; buf[nr] = 0 serves the ELL padding
  double *vtab = buf + hmax;
  uint64_t *bar = reinterpret_cast<uint64_t *>(vtab + F_MAXVALS);
  double *red = reinterpret_cast<double *>(bar + 2);
  const int tid = threadIdx.x;
  double lsx = 0.0, lcp = 0.0, lsum = 0.0;
  if (MODE == 2) { lsx = a.st->sx; lcp = a.st->cprev; }
  if (tid == 0) { mbar_init(&bar[0], 1); fence_barrier_init(); }
  if (UNI == 0)
    for (int i = tid; i < a.nvals; i += FCOL_THREADS) vtab[i] = a.vtab[i];
  __syncthreads();
  const uint32_t peer_base = mapa_u32(smem_u32(buf), me ^ 1u);
  const int64_t G = gridDim.x >> 1, pair = blockIdx.x >> 1;
  const uint32_t bytes = (uint32_t)nr * 8u;
  for (int64_t it = 0;; it++) {
    const int64_t j = pair + it * G;
    if (j >= a.ncols) break;
    if (tid == 0) {
      fence_proxy_async();
      mbar_expect_tx(&bar[0], bytes);
      const char *src = reinterpret_cast<const char *>(a.x + j * (int64_t)n + r0);
      for (uint32_t off = 0; off < bytes; off += 32768u) bulk_g2s(reinterpret_cast<char *>(buf) + off, src + off, min(32768u, bytes - off), &bar[0]);
      buf[nr] = 0.0;
    }
    mbar_wait(&bar[0], (uint32_t)(it & 1));
    cluster_sync_all();                                            // both halves of column j are in place
    double *yc = a.y + j * (int64_t)n + r0;
    for (int r = tid; r < nr; r += FCOL_THREADS) {
      uint32_t e[WT];
#pragma unroll
      for (int s = 0; s < WT; s++) e[s] = __ldg(a.ell + (size_t)s * n + r0 + r);
      double yold = 0.0;
      if (ACC) yold = __ldcs(yc + r);
      double acc = 0.0;
#pragma unroll
      for (int s = 0; s < WT; s++) {
        const int col = (int)(e[s] & F_COL_MASK);
        const uint32_t lc = (uint32_t)(col - r0);
        double xv;
        if (lc < (uint32_t)nr || col >= n) xv = buf[col >= n ? nr : (int)lc];
        else xv = ld_dsmem(peer_base + (uint32_t)(col - r0p) * 8u);
        if (UNI) acc += flip_sign(xv, e[s] & 0x80000000u);
        else acc = fma(flip_sign(vtab[(e[s] >> F_COL_BITS) & F_VID_MASK], e[s] & 0x80000000u), xv, acc);
      }
      double acc0 = UNI ? a.vuni * acc : acc;
      if (ACC) acc0 += yold;
      if (MODE == 2) {
        double *wp = a.xp + j * (int64_t)n + r0 + r;
        const double w = lsx * acc0 - lcp * __ldcs(wp);
        __stcs(wp, w);
        lsum = fma(lsx * buf[r], w, lsum);
      } else {
        __stcs(yc + r, acc0);
      }
    }
    cluster_sync_all();                                            // nobody reads this column any more
  }
  (void)nrp;
  if (MODE == 2) {
    for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
    if ((tid & 31) == 0) red[tid >> 5] = lsum;
    __syncthreads();
    if (tid < 32) {
      double r2 = red[tid];
      for (int o = 16; o > 0; o >>= 1) r2 += __shfl_xor_sync(0xffffffffu, r2, o);
      if (tid == 0) a.partials[blockIdx.x] = r2;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// pass 2: structured row-tile kernel for the single-band star geometry
// ---------------------------------------------------------------------------------------------
namespace lowtab {
__host__ __device__ constexpr int popc(int v) { int c = 0; for (; v; v &= v - 1) c++; return c; }
__host__ __device__ constexpr int binom(int n, int k) {
  if (k < 0 || k > n) return 0;
  long long r = 1;
  for (int i = 1; i <= k; i++) r = r * (n - k + i) / i;
  return (int)r;
}
// position of lo among the ascending LR-bit patterns with the same popcount
__host__ __device__ constexpr int rank(int lo) {
  int r = 0;
  const int p = popc(lo);
  for (int q = 0; q < lo; q++) if (popc(q) == p) r++;
  return r;
}
// i-th ascending LR-bit pattern with N set bits
__host__ __device__ constexpr int pat(int LR, int N, int i) {
  int r = 0;
  for (int q = 0; q < (1 << LR); q++)
    if (popc(q) == N) { if (r == i) return q; r++; }
  return -1;
}
// among the class-N patterns, the position (0-based) of pattern index i within those whose bit 0 == B
__host__ __device__ constexpr int half_index(int LR, int N, int i, int B) {
  int r = 0;
  for (int q = 0; q < i; q++) if ((pat(LR, N, q) & 1) == B) r++;
  return r;
}
__host__ __device__ constexpr int imax(int a, int b) { return a > b ? a : b; }
}  // namespace lowtab

template <typename F, int... I>
__device__ __forceinline__ void static_for_impl(F &&f, std::integer_sequence<int, I...>) {
  (f(std::integral_constant<int, I>{}), ...);
}
template <int N, typename F>
__device__ __forceinline__ void static_for(F &&f) {
  static_for_impl(f, std::make_integer_sequence<int, N>{});
}

struct SRowArgs {
  const double *x;
  double *y;
  int n;                         // DimUp (contiguous, even)
  int nf;                        // DimDw (all columns local)
  int ndw, nhigh, ngroups, nchunks, cmax;
  const int32_t *jhi;
  const uint16_t *grp;
  const int4 *chunks;
  double vk[EDGPU_MAX_SITES];    // vk[k], k = 1 .. Ns-1
  const double *diag;            // DIAG == 1: spH0d, one value per element
  const double *dr0, *dr1;       // DIAG == 2: per-row tables, dw impurity empty / occupied
  const double *dfac_s;          // DIAG == 2: per-column table (padded by 2 doubles)
  int dbg;                       // timing experiments only (wrong results): 1 = skip far hops, 2 = skip in-chunk hops, 4 = skip y read
  // sharding: this rank holds the global columns [c0, c0 + nf); x, y, diag point at local column 0
  int c0;
  int co[SROW_MAXP + 1];         // first global column of every rank
  const double *xb[SROW_MAXP];   // every rank's copy of x (peer memory over NVLink; xb[rank] == x)
};

struct SRowTile {
  const double *tl;              // shared-memory tile of this item + lane: column c at tl[(c - cb) * 32]
  const double *dsc;             // shared: dfac_s[cb ..], DIAG == 2
  double *yg;                    // y + i0 + row
  const double *dgg;             // diag + i0 + row (DIAG == 1)
  const int32_t *jhi;            // shared
  const double *vhigh;           // shared, vhigh[kk] = V_{LR+kk}
  const double *const *xb;       // shared: per-rank base of x
  size_t row;                    // row of this lane (clamped into the matrix)
  const double *x0;              // single rank: x (column 0, row 0)
  const double *tile;            // shared-memory tile of this item (column cb, row 0)
  const int *co;                 // shared: per-rank first column
  size_t n;                      // column stride in elements
  int cb, csz, nhigh, dbg;
  bool active;                   // this lane's row exists (stores only)
  double drow0, drow1;
};

// Per-group hop descriptors, built once per group by class-independent code (srow_prepare) and kept in
// shared memory (one slot per high bit and warp): where the partner group's first column is (a GENERIC
// pointer: the shared-memory tile, this GPU's L2, or a peer GPU over NVLink) and the signed hopping amplitude.
struct HopDesc {
  const double *p;
  double v;
};

// NB hops of one kind fused so that all their loads are in flight before the first use (far sources have
// ~1 us latency).  BK = the hopped bath bit is occupied in the target group: targets are the columns with the
// impurity empty, sources lo|1 in class N+1; otherwise targets have the impurity occupied, sources lo&~1 in N-1.
template <int LR, int N, int CNT, bool BK, int NB>
__device__ __forceinline__ void srow_hop_set(const HopDesc *desc, uint32_t m, size_t ss, size_t loff, double (&acc)[CNT]) {
  if constexpr ((BK && N == LR) || (!BK && N == 0)) return;        // no such targets in this class
  constexpr int HB = lowtab::imax(1, BK ? lowtab::binom(LR - 1, N) : lowtab::binom(LR - 1, N - 1));
  while (m) {
    HopDesc d[NB];
#pragma unroll
    for (int q = 0; q < NB; q++) {
      if (m == 0) { d[q].p = d[0].p; d[q].v = 0.0; continue; }     // padding slot: re-reads slot 0's source, weight 0
      d[q] = desc[__ffs((int)m) - 1];                              // warp-uniform base; this lane's row is added here
      d[q].p += loff;
      m &= m - 1;
    }
    double v[NB][HB];
    static_for<NB>([&](auto bc) {
      constexpr int q = decltype(bc)::value;
      static_for<CNT>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        constexpr int lo = lowtab::pat(LR, N, i);
        if constexpr (((lo & 1) == 0) == BK) {
          constexpr int j = lowtab::rank(BK ? (lo | 1) : (lo & ~1));
          constexpr int hi = lowtab::half_index(LR, N, i, BK ? 0 : 1);
          v[q][hi] = d[q].p[j * ss];
        }
      });
    });
    static_for<NB>([&](auto bc) {
      constexpr int q = decltype(bc)::value;
      static_for<CNT>([&](auto ic) {
        constexpr int i = decltype(ic)::value;
        constexpr int lo = lowtab::pat(LR, N, i);
        if constexpr (((lo & 1) == 0) == BK) {
          constexpr int par = lowtab::popc(lo >> 1) & 1;
          constexpr int hi = lowtab::half_index(LR, N, i, BK ? 0 : 1);
          if constexpr (par) acc[i] = fma(-d[q].v, v[q][hi], acc[i]);
          else acc[i] = fma(d[q].v, v[q][hi], acc[i]);
        }
      });
    });
  }
}

// class-independent part of a group: descriptors of its high-bit hops; returns the masks of the hops whose
// source lies in the tile (stride 32 elements) and elsewhere (stride n elements)
template <bool SH>
__device__ __forceinline__ void srow_prepare(const SRowTile &k, uint32_t h, HopDesc *desc, uint32_t &inmask, uint32_t &farmask) {
  uint32_t par = h ^ (h << 1);                                     // bit kk of par = parity of h below bit kk
  par ^= par << 2; par ^= par << 4; par ^= par << 8;
  par <<= 1;
  inmask = 0; farmask = 0;
#pragma unroll 1
  for (int kk = 0; kk < k.nhigh; kk++) {
    const int c2 = k.jhi[h ^ (1u << kk)];
    if (c2 & JHI_CUT) continue;                                    // empty group, or cut by a rank boundary (fix-up kernel)
    const int col2 = c2 & JHI_COLMASK;
    HopDesc d;
    d.v = __longlong_as_double(__double_as_longlong(k.vhigh[kk]) ^ ((long long)((par >> kk) & 1u) << 63));
    if ((unsigned)(col2 - k.cb) < (unsigned)k.csz) {
      d.p = k.tile + (col2 - k.cb) * SROW_R;
      inmask |= 1u << kk;
    } else {
      if (SH) { const int own = (c2 >> 20) & 63; d.p = k.xb[own] + (size_t)(col2 - k.co[own]) * k.n; }
      else d.p = k.x0 + (size_t)col2 * k.n;
      farmask |= 1u << kk;
    }
    if ((threadIdx.x & 31) == 0) desc[kk] = d;
  }
  __syncwarp();
}

template <int LR, int N, int DIAG, bool ACC>
__device__ __forceinline__ void srow_group(const SRowTile &k, const uint32_t h, const int base, const double (&vlow)[LR],
                                           const HopDesc *desc, uint32_t inmask, uint32_t farmask) {
  constexpr int CNT = lowtab::binom(LR, N);
  double xv[CNT], acc[CNT];
  const int lb = base - k.cb;
  const double *tl = k.tl + lb * SROW_R;
  static_for<CNT>([&](auto ic) { constexpr int i = decltype(ic)::value; xv[i] = tl[i * SROW_R]; });
  // diagonal (direct mode: factorised tables staged in shared memory); streamed inputs are read at the end
  static_for<CNT>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    constexpr int lo = lowtab::pat(LR, N, i);
    if (DIAG == 2) acc[i] = (((lo & 1) ? k.drow1 : k.drow0) + k.dsc[lb + i]) * xv[i];
    else acc[i] = 0.0;
  });
  // hops among the low bits: register to register
  static_for<CNT>([&](auto ic) {
    constexpr int i = decltype(ic)::value;
    constexpr int lo = lowtab::pat(LR, N, i);
    static_for<LR - 1>([&](auto kc) {
      constexpr int kb = decltype(kc)::value + 1;
      if constexpr (((lo >> kb) & 1) != (lo & 1)) {
        constexpr int lo2 = lo ^ (1 | (1 << kb));
        constexpr int j = lowtab::rank(lo2);
        constexpr int par = lowtab::popc(lo & ((1 << kb) - 2)) & 1;
        if constexpr (par) acc[i] = fma(-vlow[kb], xv[j], acc[i]);
        else acc[i] = fma(vlow[kb], xv[j], acc[i]);
      }
    });
  });
  // hops on the high bits: the whole group maps onto ONE other group (descriptors from srow_prepare).  Far
  // sources first (longest latency, three hops' loads in flight), then the ones in the shared-memory tile.
  if (!(k.dbg & 1)) {
    srow_hop_set<LR, N, CNT, true, SROW_NB_FAR>(desc, farmask & h, k.n, k.row, acc);
    srow_hop_set<LR, N, CNT, false, SROW_NB_FAR>(desc, farmask & ~h, k.n, k.row, acc);
  }
  if (!(k.dbg & 2)) {
    srow_hop_set<LR, N, CNT, true, SROW_NB_IN>(desc, inmask & h, SROW_R, (size_t)(threadIdx.x & 31), acc);
    srow_hop_set<LR, N, CNT, false, SROW_NB_IN>(desc, inmask & ~h, SROW_R, (size_t)(threadIdx.x & 31), acc);
  }
  double *yp = k.yg + (size_t)base * k.n;                          // yg / dgg are biased by -c0 columns
  if (DIAG == 1) {
    const double *dp = k.dgg + (size_t)base * k.n;
    double dg[CNT];
    static_for<CNT>([&](auto ic) { constexpr int i = decltype(ic)::value; dg[i] = __ldcs(dp + i * k.n); });
    static_for<CNT>([&](auto ic) { constexpr int i = decltype(ic)::value; acc[i] = fma(dg[i], xv[i], acc[i]); });
  }
  if (ACC && !(k.dbg & 4)) {
    double yold[CNT];
    static_for<CNT>([&](auto ic) { constexpr int i = decltype(ic)::value; yold[i] = __ldcs(yp + i * k.n); });
    static_for<CNT>([&](auto ic) { constexpr int i = decltype(ic)::value; acc[i] += yold[i]; });
  }
  if (k.active) static_for<CNT>([&](auto ic) { constexpr int i = decltype(ic)::value; __stcs(yp + i * k.n, acc[i]); });
}

template <int LR, int DIAG, bool ACC, int... NS>
__device__ __forceinline__ void srow_dispatch(const SRowTile &k, uint32_t h, int base, int nlow, const double (&vlow)[LR],
                                              const HopDesc *desc, uint32_t inmask, uint32_t farmask, std::integer_sequence<int, NS...>) {
  ((nlow == NS ? (srow_group<LR, NS, DIAG, ACC>(k, h, base, vlow, desc, inmask, farmask), 0) : 0), ...);
}

// 16 consumer warps + 1 producer warp.  full[b]: the TMA copies of buffer b have landed; empty[b]: every
// consumer warp is done with buffer b.  Consumer warps take the low groups of the tile from a shared
// counter and run ahead into the next buffer without a CTA-wide barrier.
template <int LR, int DIAG, bool ACC, bool SH>
__global__ void __launch_bounds__(srow_consumers(LR) + 32, 1) k_srow(const __grid_constant__ CUtensorMap tmx, SRowArgs a) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const int cpad = ((a.cmax + SROW_BC - 1) / SROW_BC) * SROW_BC;
  const int tsz = cpad * SROW_R;                                  // doubles per tile buffer
  double *tile0 = reinterpret_cast<double *>(smraw);
  double *dsc0 = tile0 + 2 * tsz;                                 // [2][cpad + 2]
  double *drw0 = dsc0 + 2 * (cpad + 2);                           // [2][2][32]
  double *vhigh = drw0 + 4 * SROW_R;                              // [32]
  uint64_t *bar = reinterpret_cast<uint64_t *>(vhigh + 32);       // full[2], empty[2]
  int *gctr = reinterpret_cast<int *>(bar + 4);                   // [2] (+2 pad)
  const double **xbs = reinterpret_cast<const double **>(gctr + 4);   // [SROW_MAXP] per-rank base of x
  int *cos = reinterpret_cast<int *>(xbs + SROW_MAXP);            // [SROW_MAXP + 1] (+ pad)
  HopDesc *desc0 = reinterpret_cast<HopDesc *>(cos + SROW_MAXP + 4);   // 16-byte aligned: every table before it is   // [24 warps][16] hop descriptors of the group in flight
  int32_t *jhi = reinterpret_cast<int32_t *>(desc0 + 24 * 16);    // [2^nhigh]
  uint16_t *grp = reinterpret_cast<uint16_t *>(jhi + (1 << a.nhigh));   // [ngroups]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  HopDesc *desc = desc0 + warp * 16;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_init(&bar[2], srow_consumers(LR) / 32);
    mbar_init(&bar[3], srow_consumers(LR) / 32);
    fence_barrier_init();
    gctr[0] = 0; gctr[1] = 0;
  }
  for (int i = tid; i < (1 << a.nhigh); i += blockDim.x) jhi[i] = a.jhi[i];
  for (int i = tid; i < a.ngroups; i += blockDim.x) grp[i] = a.grp[i];
  if (tid < 32) vhigh[tid] = (tid < a.nhigh) ? a.vk[LR + tid] : 0.0;
  if (tid < SROW_MAXP) xbs[tid] = a.xb[tid];
  if (tid <= SROW_MAXP) cos[tid] = a.co[tid];
  __syncthreads();

  const int64_t nrb = (a.n + SROW_R - 1) / SROW_R;
  const int64_t nitems = nrb * a.nchunks;
  const int64_t G = gridDim.x;
  if (warp == srow_consumers(LR) / 32) {
    // ---- producer warp ----
    for (int64_t it = 0;; it++) {
      const int64_t t = blockIdx.x + it * G;
      if (t >= nitems) break;
      const int b = (int)(it & 1);
      if (it >= 2) mbar_wait(&bar[2 + b], (uint32_t)(((it >> 1) - 1) & 1));
      const int64_t rb = t / a.nchunks;
      const int4 ch = __ldg(a.chunks + (int)(t % a.nchunks));
      const int nc = ch.w - ch.z;
      const int nops = (nc + SROW_BC - 1) / SROW_BC;
      const int64_t i0 = rb * SROW_R;
      const int nr = (int)min((int64_t)SROW_R, (int64_t)a.n - i0);
      const int lead = ch.z & 1;
      const uint32_t dsc_bytes = (uint32_t)((lead + nc + 1) & ~1) * 8u;
      if (lane == 0) {
        gctr[b] = 0;
        fence_proxy_async();
        uint32_t bytes = (uint32_t)nops * (uint32_t)(SROW_BC * SROW_R * 8);
        if (DIAG == 2) bytes += dsc_bytes + 2u * (uint32_t)nr * 8u;
        mbar_expect_tx(&bar[b], bytes);
      }
      __syncwarp();
      if (lane < nops)
        tma_load_2d(tile0 + (size_t)b * tsz + (size_t)lane * SROW_BC * SROW_R, &tmx, (int)i0, ch.z - a.c0 + lane * SROW_BC, &bar[b]);
      if (DIAG == 2) {
        if (lane == 29) bulk_g2s(dsc0 + (size_t)b * (cpad + 2), a.dfac_s + (ch.z - lead), dsc_bytes, &bar[b]);
        if (lane == 30) bulk_g2s(drw0 + (size_t)b * 2 * SROW_R, a.dr0 + i0, (uint32_t)nr * 8u, &bar[b]);
        if (lane == 31) bulk_g2s(drw0 + (size_t)b * 2 * SROW_R + SROW_R, a.dr1 + i0, (uint32_t)nr * 8u, &bar[b]);
      }
    }
    return;
  }
  // ---- consumer warps ----
  double vlow[LR];
#pragma unroll
  for (int q = 0; q < LR; q++) vlow[q] = a.vk[q];                 // vlow[0] unused
  for (int64_t it = 0;; it++) {
    const int64_t t = blockIdx.x + it * G;
    if (t >= nitems) break;
    const int b = (int)(it & 1);
    const int64_t rb = t / a.nchunks;
    const int4 ch = __ldg(a.chunks + (int)(t % a.nchunks));
    SRowTile k;
    const int64_t i0 = rb * SROW_R;
    const int64_t row = min(i0 + lane, (int64_t)a.n - 1);            // clamp: loads of a ragged last block stay in range
    k.tl = tile0 + (size_t)b * tsz + lane;
    k.dsc = dsc0 + (size_t)b * (cpad + 2) + (ch.z & 1);
    k.yg = a.y + row - (ptrdiff_t)a.c0 * a.n; k.dgg = a.diag + row - (ptrdiff_t)a.c0 * a.n;
    k.xb = xbs; k.co = cos; k.row = (size_t)row; k.x0 = a.x; k.tile = tile0 + (size_t)b * tsz;
    k.jhi = jhi; k.vhigh = vhigh;
    k.n = (size_t)a.n; k.cb = ch.z; k.csz = ch.w - ch.z; k.nhigh = a.nhigh; k.dbg = a.dbg;
    k.active = (i0 + lane) < a.n;
    mbar_wait(&bar[b], (uint32_t)((it >> 1) & 1));
    k.drow0 = 0.0; k.drow1 = 0.0;
    if (DIAG == 2) {
      k.drow0 = drw0[b * 2 * SROW_R + lane];
      k.drow1 = drw0[b * 2 * SROW_R + SROW_R + lane];
    }
    for (;;) {
      int g = 0;
      if (lane == 0) g = atomicAdd(&gctr[b], 1);
      g = __shfl_sync(0xffffffffu, g, 0) + ch.x;
      if (g >= ch.y) break;
      const uint32_t h = grp[g];
      const int base = jhi[h] & JHI_COLMASK;
      const int nlow = a.ndw - __popc(h);
      uint32_t inmask, farmask;
      srow_prepare<SH>(k, h, desc, inmask, farmask);
      srow_dispatch<LR, DIAG, ACC>(k, h, base, nlow, vlow, desc, inmask, farmask, std::make_integer_sequence<int, LR + 1>{});
      __syncwarp();                                                 // descriptors are rewritten for the next group
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar[2 + b]);                       // this warp is done with buffer b
  }
}

// ---------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------
static int pack_fast(edgpu_ctx *c, const Factor &f, FastFactor &ff) {
  std::vector<int32_t> rp((size_t)f.n + 1), cols((size_t)std::max<int64_t>(f.nnz, 1));
  std::vector<double> vals((size_t)std::max<int64_t>(f.nnz, 1));
  CK(cudaMemcpy(rp.data(), f.d_rowptr, rp.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
  if (f.nnz) {
    CK(cudaMemcpy(cols.data(), f.d_cols, (size_t)f.nnz * sizeof(int32_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(vals.data(), f.d_vals, (size_t)f.nnz * sizeof(double), cudaMemcpyDeviceToHost));
  }
  ff.n = f.n;
  ff.W = std::max(f.maxrow, 1);
  ff.WT = ff.W <= 8 ? 8 : (ff.W <= 12 ? 12 : 16);
  if (ff.W > 16) return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "factor row too long for the fast column kernel");
  std::map<double, int> ids;
  std::vector<double> vtab(1, 0.0);
  std::vector<uint32_t> ell((size_t)ff.WT * f.n, (uint32_t)f.n);       // padding: zero slot, value id 0
  for (int64_t i = 0; i < f.n; i++) {
    int k = 0;
    for (int32_t p = rp[i]; p < rp[i + 1]; p++, k++) {
      const double av = vals[p] < 0 ? -vals[p] : vals[p];
      auto it = ids.find(av);
      int id;
      if (it == ids.end()) { id = (int)vtab.size(); ids[av] = id; vtab.push_back(av); }
      else id = it->second;
      if (id >= F_MAXVALS) return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "too many distinct matrix elements for the fast column kernel");
      ell[(size_t)k * f.n + i] = (uint32_t)cols[p] | ((uint32_t)id << F_COL_BITS) | (vals[p] < 0 ? 0x80000000u : 0u);
    }
  }
  ff.nvals = (int)vtab.size();
  ff.uniform = (ff.nvals == 2);
  ff.vuni = ff.uniform ? vtab[1] : 0.0;
  CK(cudaMalloc(&ff.d_ell, ell.size() * sizeof(uint32_t)));
  CK(cudaMalloc(&ff.d_vtab, vtab.size() * sizeof(double)));
  CK(cudaMemcpy(ff.d_ell, ell.data(), ell.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ff.d_vtab, vtab.data(), vtab.size() * sizeof(double), cudaMemcpyHostToDevice));
  if (ff.uniform && f.n < 32768) {
    std::vector<uint16_t> e16(ell.size());
    for (int k = 0; k < ff.WT; k++)
      for (int64_t i = 0; i < f.n; i++) {
        const uint32_t e = ell[(size_t)k * f.n + i];
        const size_t at = (ff.WT == 8) ? (size_t)i * 8 + k : (size_t)k * f.n + i;
        e16[at] = (uint16_t)((e & 0x7FFFu) | ((e >> 31) << 15));
      }
    CK(cudaMalloc(&ff.d_ell16, e16.size() * sizeof(uint16_t)));
    CK(cudaMemcpy(ff.d_ell16, e16.data(), e16.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
  }
  return EDGPU_OK;
}

static size_t fcol_smem(int64_t n) { return (size_t)2 * (n + 2) * 8 + F_MAXVALS * 8 + 32 + 32 * 8; }
static size_t fcol2_smem(int64_t n) { return (size_t)((((n >> 1) + 1) & ~(int64_t)1) + 2) * 8 + F_MAXVALS * 8 + 16 + 32 * 8; }
static size_t srow_smem(int cmax, int nhigh, int ngroups) {
  const size_t cpad = (size_t)((cmax + SROW_BC - 1) / SROW_BC) * SROW_BC;
  size_t b = (size_t)2 * cpad * SROW_R * 8 + (size_t)2 * (cpad + 2) * 8 + 4 * SROW_R * 8 + 32 * 8 + 32 + 16 + 8 * SROW_MAXP + 4 * (SROW_MAXP + 4) +
             24 * 16 * 16 + ((size_t)4 << nhigh) + (size_t)2 * ngroups;
  return (b + 15) & ~(size_t)15;
}

// ---- pure host arithmetic of the row-kernel plan (also reachable without a GPU through
// edgpu_selftest_srow_plan, so that the CPU-only tests can check the sharding logic) -----------------
int srow_plan_host(int ns, int ndw, int64_t dimdw, int nranks, int rank, int lr, int64_t cmax_opt, SRowHostPlan &hp) {
  hp = SRowHostPlan();
  const int LR = (lr == 4 || lr == 5) ? lr : 5;
  if (ns <= LR || ns - LR > 15 || nranks > SROW_MAXP || dimdw > JHI_COLMASK) return 0;
  hp.LR = LR;
  hp.nhigh = ns - LR;
  const int P = nranks;
  hp.coloffs.assign((size_t)P + 1, 0);
  for (int p = 0; p <= P; p++) {
    int64_t q = 0, off = dimdw;
    if (p < P) edgpu_split(dimdw, P, p, &q, &off);
    hp.coloffs[(size_t)p] = (int)off;
  }
  auto owner_of = [&](int col) { int p = 0; while (p + 1 < P && col >= hp.coloffs[(size_t)p + 1]) p++; return p; };
  const int c0 = hp.coloffs[(size_t)rank], c1 = hp.coloffs[(size_t)rank + 1];
  const int nh = 1 << hp.nhigh;
  hp.jhi.assign((size_t)nh, -1);
  int64_t col = 0;
  for (int h = 0; h < nh; h++) {
    const int nlow = ndw - __builtin_popcount((unsigned)h);
    if (nlow < 0 || nlow > LR) continue;
    const int sz = lowtab::binom(LR, nlow);
    const int own = owner_of((int)col);
    const bool cut = owner_of((int)col + sz - 1) != own;           // the group is split between two ranks
    hp.jhi[(size_t)h] = (int32_t)col | (own << 20) | (cut ? JHI_CUT : 0);
    hp.grp.push_back((uint16_t)h);
    hp.gsize.push_back(sz); hp.gbase.push_back((int)col); hp.gcut.push_back(cut ? 1 : 0);
    col += sz;
  }
  if (col != dimdw) return -1;                                     // the Lin table must cover the basis exactly
  const int ngroups = (int)hp.grp.size();
  // groups that lie entirely on this rank: a contiguous run [g0, g1) of the global list
  int g0 = 0, g1 = 0;
  {
    int g = 0;
    while (g < ngroups && (hp.gbase[(size_t)g] < c0 || hp.gcut[(size_t)g])) { if (hp.gbase[(size_t)g] >= c1) break; g++; }
    g0 = g;
    while (g < ngroups && !hp.gcut[(size_t)g] && hp.gbase[(size_t)g] + hp.gsize[(size_t)g] <= c1) g++;
    g1 = g;
    if (g0 < ngroups && hp.gbase[(size_t)g0] >= c1) g1 = g0;       // nothing whole on this rank
  }
  hp.g0 = g0; hp.g1 = g1;
  // chunk size from the shared-memory budget (or the option), chunks = runs of whole groups
  int cmax = (int)((SMEM_LIMIT - 12288 - ((size_t)4 << hp.nhigh) - 2 * (size_t)ngroups) / (2 * (SROW_R + 1) * 8));
  cmax = cmax / SROW_BC * SROW_BC;
  if (cmax > 32 * SROW_BC) cmax = 32 * SROW_BC;             // one tensor copy per lane of the producer warp
  // measured on B200 (C3, chunk sweep 96..352): tiles of ~160 columns beat the largest that fits by 25 %;
  // the L1 that is left over (228 KB - shared memory) serves the out-of-tile sources
  if (cmax_opt <= 0 && cmax > 160) cmax = 160;
  if (cmax_opt > 0 && cmax_opt < cmax) cmax = (int)cmax_opt;
  if (cmax < lowtab::binom(LR, LR / 2)) return 0;
  // balance: all chunks about the same size
  const int nloccols = (g1 > g0) ? hp.gbase[(size_t)g1 - 1] + hp.gsize[(size_t)g1 - 1] - hp.gbase[(size_t)g0] : 0;
  const int nch0 = std::max(1, (nloccols + cmax - 1) / cmax);
  const int target = (nloccols + nch0 - 1) / nch0;
  int gb = g0, cb = (g1 > g0) ? hp.gbase[(size_t)g0] : 0, cur = 0;
  for (int g = g0; g < g1; g++) {
    if (cur > 0 && (cur + hp.gsize[(size_t)g] > cmax || cur >= target)) {
      hp.chunks.push_back(make_int4(gb, g, cb, cb + cur));
      gb = g; cb += cur; cur = 0;
    }
    cur += hp.gsize[(size_t)g];
  }
  if (cur > 0) hp.chunks.push_back(make_int4(gb, g1, cb, cb + cur));
  hp.cmax = lowtab::binom(LR, LR / 2);
  for (auto &ch : hp.chunks) hp.cmax = std::max(hp.cmax, ch.w - ch.z);
  hp.smem = srow_smem(hp.cmax, hp.nhigh, ngroups);
  if (hp.smem > SMEM_LIMIT) return 0;
  hp.ok = true;
  return 1;
}

// fix-up list: every dw hop (target <- source) of a LOCAL target column for which the target's or the source's
// low group is cut by a rank boundary (the structured kernel skips exactly those), from the reference-order CSR
// of spH0dws.  Targets in cut groups are computed entirely by the fix-up kernel (diagonal included).
void srow_fix_host(SRowHostPlan &hp, int rank, int64_t dimdw, const int32_t *rp, const int32_t *cc, const double *vv) {
  const int P = (int)hp.coloffs.size() - 1;
  hp.tptr.assign(1, 0); hp.tcol.clear(); hp.tinit.clear(); hp.eown.clear(); hp.esrc.clear(); hp.eval.clear();
  if (P <= 1) return;
  auto owner_of = [&](int col) { int p = 0; while (p + 1 < P && col >= hp.coloffs[(size_t)p + 1]) p++; return p; };
  const int c0 = hp.coloffs[(size_t)rank], c1 = hp.coloffs[(size_t)rank + 1];
  std::vector<char> colcut((size_t)dimdw, 0);
  for (size_t g = 0; g < hp.grp.size(); g++)
    if (hp.gcut[g]) for (int i = 0; i < hp.gsize[g]; i++) colcut[(size_t)hp.gbase[g] + i] = 1;
  for (int t = c0; t < c1; t++) {
    const bool tc = colcut[(size_t)t] != 0;
    const size_t before = hp.eown.size();
    for (int32_t q = rp[(size_t)t]; q < rp[(size_t)t + 1]; q++) {
      const int sc = cc[(size_t)q];
      if (!tc && !colcut[(size_t)sc]) continue;
      const int own = owner_of(sc);
      hp.eown.push_back(own); hp.esrc.push_back(sc - hp.coloffs[(size_t)own]); hp.eval.push_back(vv[(size_t)q]);
    }
    if (tc || hp.eown.size() > before) { hp.tcol.push_back(t - c0); hp.tinit.push_back(tc ? 1 : 0); hp.tptr.push_back((int)hp.eown.size()); }
  }
}

static int build_srow(edgpu_ctx *c, SRowPlan &sr, int LR) {
  // single band, star geometry, no inter-orbital terms: every dw hop is bit 0 <-> bit k
  sr.ok = false;
  if (c->dp.norb != 1 || c->dp.jhflag || (c->dimup & 1)) return EDGPU_OK;
  if (c->opt_srow_lr == 4 || c->opt_srow_lr == 5) LR = (int)c->opt_srow_lr;
  SRowHostPlan hp;
  const int prc = srow_plan_host(c->ns, c->ndw, c->dimdw, c->nranks, c->rank, LR, c->opt_srow_cmax, hp);
  if (prc < 0) return edgpu_set_err(EDGPU_ERR_INVALID, "internal: Lin table does not cover the dw basis");
  if (prc == 0) return EDGPU_OK;
  sr.LR = hp.LR; sr.nhigh = hp.nhigh; sr.ngroups = (int)hp.grp.size(); sr.nchunks = (int)hp.chunks.size();
  sr.cmax = hp.cmax; sr.smem = hp.smem;
  for (size_t p = 0; p < hp.coloffs.size(); p++) sr.coloffs[p] = hp.coloffs[p];
  for (int k = 0; k < EDGPU_MAX_SITES; k++) sr.vk[k] = 0.0;
  for (int k = 1; k < c->ns; k++) sr.vk[k] = c->dp.bv_dw[k - 1];
  std::vector<int4> chunks = hp.chunks;
  if (chunks.empty()) chunks.push_back(make_int4(0, 0, 0, 0));
  CK(cudaMalloc(&sr.d_jhi, hp.jhi.size() * sizeof(int32_t)));
  CK(cudaMalloc(&sr.d_grp, std::max<size_t>(hp.grp.size(), 1) * sizeof(uint16_t)));
  CK(cudaMalloc(&sr.d_chunks, chunks.size() * sizeof(int4)));
  CK(cudaMemcpy(sr.d_jhi, hp.jhi.data(), hp.jhi.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(sr.d_grp, hp.grp.data(), hp.grp.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(sr.d_chunks, chunks.data(), chunks.size() * sizeof(int4), cudaMemcpyHostToDevice));
  if (c->up.d_dfac) {                                              // direct mode: per-row diagonal tables
    std::vector<double> d0((size_t)c->dimup + 32, 0.0), d1((size_t)c->dimup + 32, 0.0);
    std::vector<int32_t> mu((size_t)c->dimup);
    CK(cudaMemcpy(d0.data(), c->up.d_dfac, (size_t)c->dimup * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(mu.data(), c->up.d_map, (size_t)c->dimup * sizeof(int32_t), cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < c->dimup; i++) d1[(size_t)i] = d0[(size_t)i] + ((mu[(size_t)i] & 1) ? c->dp.uloc[0] : 0.0);
    CK(cudaMalloc(&sr.d_dr0, d0.size() * sizeof(double)));
    CK(cudaMalloc(&sr.d_dr1, d1.size() * sizeof(double)));
    CK(cudaMemcpy(sr.d_dr0, d0.data(), d0.size() * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(sr.d_dr1, d1.data(), d1.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  sr.nfix = 0;
  if (c->nranks > 1) {
    std::vector<int32_t> rp((size_t)c->dw.n + 1), cc((size_t)std::max<int64_t>(c->dw.nnz, 1));
    std::vector<double> vv((size_t)std::max<int64_t>(c->dw.nnz, 1));
    CK(cudaMemcpy(rp.data(), c->dw.d_rowptr, rp.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (c->dw.nnz) {
      CK(cudaMemcpy(cc.data(), c->dw.d_cols, (size_t)c->dw.nnz * sizeof(int32_t), cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(vv.data(), c->dw.d_vals, (size_t)c->dw.nnz * sizeof(double), cudaMemcpyDeviceToHost));
    }
    srow_fix_host(hp, c->rank, c->dimdw, rp.data(), cc.data(), vv.data());
    sr.nfix = (int)hp.tcol.size();
    if (sr.nfix) {
      const size_t ne = std::max<size_t>(hp.eown.size(), 1);
      hp.eown.resize(ne); hp.esrc.resize(ne); hp.eval.resize(ne);
      CK(cudaMalloc(&sr.d_ftptr, hp.tptr.size() * sizeof(int)));
      CK(cudaMalloc(&sr.d_ftcol, hp.tcol.size() * sizeof(int)));
      CK(cudaMalloc(&sr.d_ftinit, hp.tinit.size()));
      CK(cudaMalloc(&sr.d_feown, ne * sizeof(int)));
      CK(cudaMalloc(&sr.d_fesrc, ne * sizeof(int)));
      CK(cudaMalloc(&sr.d_feval, ne * sizeof(double)));
      CK(cudaMemcpy(sr.d_ftptr, hp.tptr.data(), hp.tptr.size() * sizeof(int), cudaMemcpyHostToDevice));
      CK(cudaMemcpy(sr.d_ftcol, hp.tcol.data(), hp.tcol.size() * sizeof(int), cudaMemcpyHostToDevice));
      CK(cudaMemcpy(sr.d_ftinit, hp.tinit.data(), hp.tinit.size(), cudaMemcpyHostToDevice));
      CK(cudaMemcpy(sr.d_feown, hp.eown.data(), ne * sizeof(int), cudaMemcpyHostToDevice));
      CK(cudaMemcpy(sr.d_fesrc, hp.esrc.data(), ne * sizeof(int), cudaMemcpyHostToDevice));
      CK(cudaMemcpy(sr.d_feval, hp.eval.data(), ne * sizeof(double), cudaMemcpyHostToDevice));
    }
  }
  sr.ok = true;
  return EDGPU_OK;
}

int fast_plan_free(edgpu_ctx *c) {
  if (!c->fplan) return EDGPU_OK;
  for (int k = 0; k < 2; k++) { cudaFree(c->fplan->ff[k].d_ell); cudaFree(c->fplan->ff[k].d_ell16); cudaFree(c->fplan->ff[k].d_vtab); }
  cudaFree(c->fplan->sr.d_jhi); cudaFree(c->fplan->sr.d_grp); cudaFree(c->fplan->sr.d_chunks);
  cudaFree(c->fplan->sr.d_dr0); cudaFree(c->fplan->sr.d_dr1);
  { SRowPlan &r = c->fplan->sr; cudaFree(r.d_ftptr); cudaFree(r.d_ftcol); cudaFree(r.d_ftinit); cudaFree(r.d_feown); cudaFree(r.d_fesrc); cudaFree(r.d_feval); }
  delete c->fplan;
  c->fplan = nullptr;
  return EDGPU_OK;
}

template <int WT, int DIAG>
static cudaError_t set_fcol_attr() {
  cudaError_t e = cudaSuccess;
#define SETF(U, A) if (e == cudaSuccess) e = cudaFuncSetAttribute(k_fcol<WT, DIAG, U, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT)
  SETF(0, 0); SETF(1, 0); SETF(2, 0); SETF(0, 1); SETF(1, 1); SETF(2, 1);
  if (DIAG == 0) { SETF(0, 2); SETF(1, 2); SETF(2, 2); }
#undef SETF
  return e;
}
template <int LR, int DIAG>
static cudaError_t set_srow_attr() {
  cudaError_t e = cudaFuncSetAttribute(k_srow<LR, DIAG, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_srow<LR, DIAG, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_srow<LR, DIAG, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_srow<LR, DIAG, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
  return e;
}

int fast_plan_build(edgpu_ctx *c) {
  if (c->fplan) return EDGPU_OK;
  static bool attr_done = false;
  if (!attr_done) {
    CK((set_fcol_attr<8, 0>())); CK((set_fcol_attr<8, 1>())); CK((set_fcol_attr<8, 2>()));
    CK((set_fcol_attr<12, 0>())); CK((set_fcol_attr<12, 1>())); CK((set_fcol_attr<12, 2>()));
    CK((set_fcol_attr<16, 0>())); CK((set_fcol_attr<16, 1>())); CK((set_fcol_attr<16, 2>()));
    CK((set_srow_attr<4, 0>())); CK((set_srow_attr<4, 1>())); CK((set_srow_attr<4, 2>()));
    CK((set_srow_attr<5, 0>())); CK((set_srow_attr<5, 1>())); CK((set_srow_attr<5, 2>()));
    attr_done = true;
  }
  FastPlan *p = new FastPlan();
  c->fplan = p;
  {
    static bool attr2_done = false;
    if (!attr2_done) {
#define SET2(W, U, M) CK(cudaFuncSetAttribute(k_fcol2<W, U, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT))
      SET2(8, 0, 0); SET2(8, 1, 0); SET2(8, 0, 1); SET2(8, 1, 1); SET2(8, 0, 2); SET2(8, 1, 2);
      SET2(12, 0, 0); SET2(12, 1, 0); SET2(12, 0, 1); SET2(12, 1, 1); SET2(12, 0, 2); SET2(12, 1, 2);
      SET2(16, 0, 0); SET2(16, 1, 0); SET2(16, 0, 1); SET2(16, 1, 1); SET2(16, 0, 2); SET2(16, 1, 2);
#undef SET2
      attr2_done = true;
    }
  }
  for (int k = 0; k < 2; k++) {
    const Factor &f = k ? c->dw : c->up;
    p->col_smem[k] = fcol_smem(f.n);
    p->col_ok[k] = (f.n % 2 == 0) && f.n < (1 << F_COL_BITS) && p->col_smem[k] <= SMEM_LIMIT && f.maxrow <= 16;
    p->col2_smem[k] = fcol2_smem(f.n);
    p->col2_ok[k] = (f.n % 2 == 0) && f.n >= 8 && f.n < (1 << F_COL_BITS) && p->col2_smem[k] <= SMEM_LIMIT && f.maxrow <= 16 && c->sm_count >= 2;
    if (p->col_ok[k] || p->col2_ok[k]) {
      int rc = pack_fast(c, f, p->ff[k]);
      if (rc == EDGPU_ERR_UNSUPPORTED) { p->col_ok[k] = false; p->col2_ok[k] = false; }
      else if (rc) { fast_plan_free(c); return rc; }
    }
  }
  int rc = build_srow(c, p->sr, 5);
  if (rc) { fast_plan_free(c); return rc; }
  return EDGPU_OK;
}

bool fast_supported_local(edgpu_ctx *c) {
  if (!c->hstatus || c->dp.jhflag) return false;
  if (fast_plan_build(c)) return false;
  return (c->fplan->col_ok[0] || c->fplan->col2_ok[0]) && c->fplan->sr.ok;
}
bool fast_supported_col(edgpu_ctx *c, int k) {
  if (!c->hstatus || c->dp.jhflag) return false;
  if (fast_plan_build(c)) return false;
  return c->fplan->col_ok[k] || c->fplan->col2_ok[k];
}

template <int WT, int DIAG, int MODE>
static void launch_fcol_m(int uni, int grid, size_t smem, cudaStream_t st, const FColArgs &a) {
  if (uni == 2) k_fcol<WT, DIAG, 2, MODE><<<grid, FCOL_THREADS, smem, st>>>(a);
  else if (uni == 1) k_fcol<WT, DIAG, 1, MODE><<<grid, FCOL_THREADS, smem, st>>>(a);
  else k_fcol<WT, DIAG, 0, MODE><<<grid, FCOL_THREADS, smem, st>>>(a);
}
template <int WT, int DIAG>
static void launch_fcol(int uni, int mode, int grid, size_t smem, cudaStream_t st, const FColArgs &a) {
  if (mode == 2) { if constexpr (DIAG == 0) launch_fcol_m<WT, 0, 2>(uni, grid, smem, st, a); }
  else if (mode == 1) launch_fcol_m<WT, DIAG, 1>(uni, grid, smem, st, a);
  else launch_fcol_m<WT, DIAG, 0>(uni, grid, smem, st, a);
}
template <int DIAG>
static void launch_fcol_w(int WT, int uni, int mode, int grid, size_t smem, cudaStream_t st, const FColArgs &a) {
  if (WT == 8) launch_fcol<8, DIAG>(uni, mode, grid, smem, st, a);
  else if (WT == 12) launch_fcol<12, DIAG>(uni, mode, grid, smem, st, a);
  else launch_fcol<16, DIAG>(uni, mode, grid, smem, st, a);
}

template <int WT>
static void launch_fcol2_w(int uni, int mode, int grid, size_t smem, cudaStream_t st, const FColArgs &a) {
  if (uni) {
    if (mode == 2) k_fcol2<WT, 1, 2><<<grid, FCOL_THREADS, smem, st>>>(a);
    else if (mode == 1) k_fcol2<WT, 1, 1><<<grid, FCOL_THREADS, smem, st>>>(a);
    else k_fcol2<WT, 1, 0><<<grid, FCOL_THREADS, smem, st>>>(a);
  } else {
    if (mode == 2) k_fcol2<WT, 0, 2><<<grid, FCOL_THREADS, smem, st>>>(a);
    else if (mode == 1) k_fcol2<WT, 0, 1><<<grid, FCOL_THREADS, smem, st>>>(a);
    else k_fcol2<WT, 0, 0><<<grid, FCOL_THREADS, smem, st>>>(a);
  }
}
static void launch_fcol2(int WT, int uni, int mode, int grid, size_t smem, cudaStream_t st, const FColArgs &a) {
  if (WT == 8) launch_fcol2_w<8>(uni, mode, grid, smem, st, a);
  else if (WT == 12) launch_fcol2_w<12>(uni, mode, grid, smem, st, a);
  else launch_fcol2_w<16>(uni, mode, grid, smem, st, a);
}

// y (+)= [Hd o x +] F_k x on a matrix whose contiguous dimension is factor k's index
int fast_apply_col(edgpu_ctx *c, int k, bool with_diag, bool acc, const double *d_x, double *d_y, int64_t ncols, int64_t coloff,
                   double *d_xp, int *npartials) {
  TRY(fast_plan_build(c));
  FastPlan *p = c->fplan;
  const bool use2 = p->col2_ok[k] && (!p->col_ok[k] || c->opt_col_cluster);
  if (!p->col_ok[k] && !use2) return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "fast column kernel does not cover this factor");
  if ((reinterpret_cast<uintptr_t>(d_x) & 15) != 0) return edgpu_set_err(EDGPU_ERR_INVALID, "fast H*v needs 16-byte aligned vectors");
  const FastFactor &ff = p->ff[k];
  FColArgs a{};
  a.x = d_x; a.y = d_y; a.n = (int)ff.n; a.ncols = ncols; a.coloff = coloff;
  a.ell = ff.d_ell; a.ell16 = ff.d_ell16; a.vtab = ff.d_vtab; a.nvals = ff.nvals; a.vuni = ff.vuni;
  a.diag = c->d_diag;
  const Factor &fc = k ? c->dw : c->up, &fs = k ? c->up : c->dw;
  a.dfac_c = fc.d_dfac; a.dfac_s = fs.d_dfac; a.map_c = fc.d_map; a.map_s = fs.d_map;
  a.norb = c->dp.norb; a.ust = c->dp.ust;
  for (int i = 0; i < EDGPU_MAX_ORB; i++) a.uloc[i] = c->dp.uloc[i];
  const int grid = (int)std::min<int64_t>(ncols, c->sm_count);
  if (grid < 1) return EDGPU_OK;
  const int diag = !with_diag ? 0 : (c->d_diag ? 1 : 2);
  const int mode = d_xp ? 2 : (acc ? 1 : 0);
  if (mode == 2 && (diag != 0 || !acc)) return edgpu_set_err(EDGPU_ERR_INVALID, "Lanczos epilogue needs the accumulate form without diagonal");
  a.xp = d_xp; a.st = c->d_st; a.partials = c->d_partials;
  if (npartials) *npartials = grid;
  if (use2) {
    if (diag != 0) return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "cluster column kernel has no fused diagonal");
    const int pairs = (int)std::min<int64_t>(ncols, c->sm_count / 2);
    const int uni2 = (ff.uniform && c->opt_no_uniform != 1) ? 1 : 0;
    if (npartials) *npartials = 2 * pairs;
    launch_fcol2(ff.WT, uni2, mode, 2 * pairs, p->col2_smem[k], c->stream, a);
    CKL(c);
    return EDGPU_OK;
  }
  // no_uniform: 0 = best available, 1 = force the value-table kernel, 2 = uniform kernel with 4-byte entries
  int uni = 0;
  if (ff.uniform && c->opt_no_uniform != 1) uni = (ff.d_ell16 && c->opt_no_uniform != 2) ? 2 : 1;
  if (diag == 0) launch_fcol_w<0>(ff.WT, uni, mode, grid, p->col_smem[k], c->stream, a);
  else if (diag == 1) launch_fcol_w<1>(ff.WT, uni, mode, grid, p->col_smem[k], c->stream, a);
  else launch_fcol_w<2>(ff.WT, uni, mode, grid, p->col_smem[k], c->stream, a);
  CKL(c);
  return EDGPU_OK;
}

template <int LR, bool SH>
static void launch_srow_s(int diag, bool acc, int grid, size_t smem, cudaStream_t st, const CUtensorMap &tmx, const SRowArgs &a) {
  const int nt = srow_consumers(LR) + 32;
  if (acc) {
    if (diag == 0) k_srow<LR, 0, true, SH><<<grid, nt, smem, st>>>(tmx, a);
    else if (diag == 1) k_srow<LR, 1, true, SH><<<grid, nt, smem, st>>>(tmx, a);
    else k_srow<LR, 2, true, SH><<<grid, nt, smem, st>>>(tmx, a);
  } else {
    if (diag == 0) k_srow<LR, 0, false, SH><<<grid, nt, smem, st>>>(tmx, a);
    else if (diag == 1) k_srow<LR, 1, false, SH><<<grid, nt, smem, st>>>(tmx, a);
    else k_srow<LR, 2, false, SH><<<grid, nt, smem, st>>>(tmx, a);
  }
}
template <int LR>
static void launch_srow(bool sharded, int diag, bool acc, int grid, size_t smem, cudaStream_t st, const CUtensorMap &tmx, const SRowArgs &a) {
  if (sharded) launch_srow_s<LR, true>(diag, acc, grid, smem, st, tmx, a);
  else launch_srow_s<LR, false>(diag, acc, grid, smem, st, tmx, a);
}

// 2-D tensor map of a column-major double matrix (n0 contiguous), box = b0 x b1 elements
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int make_tmap_2d(CUtensorMap *tm, const double *base, uint64_t n0, uint64_t n1, uint32_t b0, uint32_t b1) {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) return edgpu_set_err(EDGPU_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    fn = (PFN_encodeTiled)p;
  }
  const cuuint64_t dims[2] = {n0, n1};
  const cuuint64_t strides[1] = {n0 * 8};
  const cuuint32_t box[2] = {b0, b1};
  const cuuint32_t es[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(base), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return edgpu_set_err(EDGPU_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return EDGPU_OK;
}

// fix-up for the sharded vector (see build_srow): one CTA per listed local target column
struct SFixArgs {
  const double *x;               // local x
  double *y;
  int n, c0;
  const int *tptr, *tcol, *eown, *esrc;
  const unsigned char *tinit;
  const double *eval;
  const double *xb[SROW_MAXP];
  const double *diag;            // stored: local spH0d
  const double *dr0, *dr1, *dfac_s;
  const int32_t *map_s;
  int diagmode, acc;
};
__global__ void __launch_bounds__(256) k_sfix(SFixArgs a) {
  const int t = blockIdx.x;
  const int tl = a.tcol[t], e0 = a.tptr[t], e1 = a.tptr[t + 1];
  const bool init = a.tinit[t] != 0;
  const size_t off = (size_t)tl * a.n;
  double dcol = 0.0;
  bool nd = false;
  if (init && a.diagmode == 2) { dcol = a.dfac_s[a.c0 + tl]; nd = (a.map_s[a.c0 + tl] & 1) != 0; }
  for (int r = blockIdx.y * blockDim.x + threadIdx.x; r < a.n; r += gridDim.y * blockDim.x) {
    double acc;
    if (init) {
      double d = 0.0;
      if (a.diagmode == 1) d = a.diag[off + r];
      if (a.diagmode == 2) d = (nd ? a.dr1[r] : a.dr0[r]) + dcol;
      acc = d * a.x[off + r];
      if (a.acc) acc += a.y[off + r];
    } else {
      acc = a.y[off + r];                                           // written by k_srow just before
    }
    for (int e = e0; e < e1; e++) acc = fma(a.eval[e], a.xb[a.eown[e]][(size_t)a.esrc[e] * a.n + r], acc);
    a.y[off + r] = acc;
  }
}

// y (+)= [Hd o x +] x Hdw^T on the local shard.  nranks == 1: every column is local.  nranks > 1: sources on
// other ranks are read from the peers' copies of x (xpeer[p] = rank p's pointer for the same vector).
int fast_apply_row(edgpu_ctx *c, bool with_diag, bool acc, const double *d_x, double *d_y, const double *const *xpeer) {
  TRY(fast_plan_build(c));
  const SRowPlan &sr = c->fplan->sr;
  if (!sr.ok) return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "structured row kernel does not cover this model");
  if ((reinterpret_cast<uintptr_t>(d_x) & 15) != 0) return edgpu_set_err(EDGPU_ERR_INVALID, "fast H*v needs 16-byte aligned vectors");
  if (c->nranks > 1 && !xpeer) return edgpu_set_err(EDGPU_ERR_INVALID, "sharded row kernel needs the peers' vectors");
  SRowArgs a{};
  a.x = d_x; a.y = d_y; a.n = (int)c->dimup; a.nf = (int)c->qdw;
  a.ndw = c->ndw; a.nhigh = sr.nhigh; a.ngroups = sr.ngroups; a.nchunks = sr.nchunks; a.cmax = sr.cmax;
  a.jhi = sr.d_jhi; a.grp = sr.d_grp; a.chunks = sr.d_chunks;
  for (int k = 0; k < EDGPU_MAX_SITES; k++) a.vk[k] = sr.vk[k];
  a.dbg = (int)c->opt_dbg;
  a.diag = c->d_diag; a.dr0 = sr.d_dr0; a.dr1 = sr.d_dr1; a.dfac_s = c->dw.d_dfac;
  a.c0 = (int)c->coloff;
  for (int p = 0; p <= SROW_MAXP; p++) a.co[p] = sr.coloffs[p < c->nranks ? p : c->nranks];
  for (int p = 0; p < SROW_MAXP; p++) a.xb[p] = (c->nranks == 1 || p >= c->nranks) ? d_x : xpeer[p];
  a.xb[c->rank] = d_x;
  const int diag = !with_diag ? 0 : (c->d_diag ? 1 : 2);
  const int64_t nitems = ((c->dimup + SROW_R - 1) / SROW_R) * sr.nchunks;
  if (sr.nchunks > 0 && c->fplan->sr.cmax > 0 && nitems > 0 && !(sr.nchunks == 1 && c->qdw == 0)) {
    const int grid = (int)std::min<int64_t>(nitems, c->sm_count);
    CUtensorMap tmx;
    TRY(make_tmap_2d(&tmx, d_x, (uint64_t)c->dimup, (uint64_t)c->qdw, SROW_R, SROW_BC));
    // a rank may own no whole group at all (tiny sectors): the chunk table then holds one empty chunk
    if (sr.LR == 4) launch_srow<4>(c->nranks > 1, diag, acc, grid, sr.smem, c->stream, tmx, a);
    else launch_srow<5>(c->nranks > 1, diag, acc, grid, sr.smem, c->stream, tmx, a);
    CKL(c);
  }
  if (sr.nfix > 0) {
    SFixArgs f{};
    f.x = d_x; f.y = d_y; f.n = (int)c->dimup; f.c0 = (int)c->coloff;
    f.tptr = sr.d_ftptr; f.tcol = sr.d_ftcol; f.eown = sr.d_feown; f.esrc = sr.d_fesrc; f.tinit = sr.d_ftinit; f.eval = sr.d_feval;
    for (int p = 0; p < SROW_MAXP; p++) f.xb[p] = a.xb[p];
    f.diag = c->d_diag; f.dr0 = sr.d_dr0; f.dr1 = sr.d_dr1; f.dfac_s = c->dw.d_dfac; f.map_s = c->dw.d_map;
    f.diagmode = diag; f.acc = acc ? 1 : 0;
    dim3 grid((unsigned)sr.nfix, (unsigned)std::max<int64_t>(1, std::min<int64_t>(32, (c->dimup + 1023) / 1024)));
    k_sfix<<<grid, 256, 0, c->stream>>>(f);
    CKL(c);
  }
  return EDGPU_OK;
}

// peers' pointers of a vector that lives in the symmetric slab (nullptr if it does not)
static const double *const *peer_ptrs(edgpu_ctx *c, const double *d_x, const double **buf) {
  if (c->nranks == 1) return nullptr;
  const int64_t off = sym_offset(c, d_x);
  if (off < 0) return nullptr;
  for (int p = 0; p < c->nranks; p++) buf[p] = reinterpret_cast<const double *>(c->sym_peer[p] + off);
  return buf;
}
bool fast_peer_ready(edgpu_ctx *c, const double *d_x) {
  return c->nranks > 1 && c->nranks <= SROW_MAXP && c->sym_ok && sym_offset(c, d_x) >= 0;
}

// d_xp != nullptr: Lanczos form -- d_y only holds the row-pass partial result, w = sx*(H x) - cprev*xp goes to
// d_xp and the per-CTA partial sums of (sx*x).w to c->d_partials (*npartials of them)
int fast_apply_local(edgpu_ctx *c, const double *d_x, double *d_y, double *d_xp, int *npartials) {
  if (c->opt_dbg & 8) {                                            // experiment: column pass first
    prof_mark(c, "k_fcol");
    TRY(fast_apply_col(c, 0, false, false, d_x, d_y, c->qdw, c->coloff, nullptr, nullptr));
    prof_mark(c, "k_srow");
    TRY(fast_apply_row(c, true, true, d_x, d_y, nullptr));
    return EDGPU_OK;
  }
  // row tiles first (y = Hd o x + x Hdw^T, write only), then whole columns (y += Hup x, contiguous RMW)
  const double *pb[64];
  const double *const *xpeer = peer_ptrs(c, d_x, pb);
  if (c->nranks > 1 && !xpeer) return edgpu_set_err(EDGPU_ERR_INVALID, "sharded fast H*v: the vector is not in the symmetric slab");
  // Peers read x while this rank reads theirs.  One stream-ordered barrier per application orders every rank's
  // earlier writes of x before the reads (RAW) and, because each rank enqueues it after its previous H*v, every
  // rank's previous remote reads before anybody's later overwrite of those buffers (WAR).
  if (c->nranks > 1) TRY(comm_barrier(c));
  prof_mark(c, "k_srow");
  TRY(fast_apply_row(c, true, false, d_x, d_y, xpeer));
  prof_mark(c, "k_fcol");
  TRY(fast_apply_col(c, 0, false, true, d_x, d_y, c->qdw, c->coloff, d_xp, npartials));
  return EDGPU_OK;
}
