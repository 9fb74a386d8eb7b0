// hxv_fast.cu -- the engine's fast H*v path: two HBM passes, each staged through shared memory by the TMA
// engine (cp.async.bulk / cp.async.bulk.tensor + mbarrier, multi-buffered, persistent CTAs, one CTA per SM,
// a dedicated producer warp and full/empty barriers instead of CTA-wide barriers).
//
//   y = Hd o x + Hup x + x Hdw^T       x(i_up, i_dw) column-major, i_up contiguous
//
//   pass 1  k_srow : y = Hd o x + x Hdw^T (write only).  Tile = 32 consecutive i_up rows x one ALIGNED chunk of
//                    i_dw columns (2-D tensor-map TMA boxes); lanes run along i_up so every shared-memory access
//                    is unit stride, and the dw hops are generated from the bit structure of the star geometry
//                    (Norb = 1: every hop is impurity bit 0 <-> bath bit k).  The columns of one "low group"
//                    (same high word h, LR low bits) live in registers; hops among the low bits are register to
//                    register; a hop on a high bit moves the whole group onto ONE other group.  A chunk is the set
//                    of groups that share the top bits of h, so it is closed under the T lowest high bits: those
//                    hops read the shared-memory tile, only the hops on the remaining top bits read L2.  Everything
//                    about a group (tile offset, class, parities, partner columns) is precomputed once per sector
//                    in an 80-byte record that the producer warp bulk-copies next to the tile.
//   pass 2  k_fcol : y += F x(:, j) with one WHOLE column per stage in shared memory (a contiguous 8*DimUp byte
//                    bulk copy).  F is any one-spin factor (spH0ups / spH0dws, ED_HAMILTONIAN/stored/H_up.f90,
//                    H_dw.f90) in a packed ELL form streamed from L2; every gather hits shared memory.  MODE 2
//                    fuses the first Lanczos vector update (w = s*Hx - c*x_prev, alpha partials) into the epilogue.
//           k_fcol2: the same for columns that exceed one SM (Ns = 18): a 2-CTA cluster holds the column, the
//                    other half is read through distributed shared memory.
//
// Sharded vector (i_dw columns split over the GPUs of one NVLink domain, ED_HAMILTONIAN.f90:96-110): k_srow only
// touches local memory.  The dw hops whose source column lives on another rank -- and the few that touch a low
// group cut by a rank boundary -- come from per-column source lists.  The OWNER of a source column stores it into
// the halo buffer of every rank that lists it (k_halo_push: posted NVLink writes into the peer-mapped symmetric
// slab of comm.cu, on a second stream and on its own SMs while k_srow works; nobody reads remote memory), raises
// an arrival flag on each peer, and pass 2 adds the arrived columns (amplitude * column) while it adds y.  Two halo
// buffers alternate, so no cross-rank barrier is needed at all.
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
#define SROW_R 32                 // rows per tile = lanes of a warp
#define SROW_BC 32                // columns per TMA box (32 rows x 32 columns x 8 B = 8 KB per copy)
#define SROW_THREADS 512          // 15 consumer warps + 1 producer warp, 128 registers each
#define SROW_CONSUMERS (SROW_THREADS - 32)
#ifndef SROW_STAGES
#define SROW_STAGES 2
#endif
#ifndef SROW_NB_FAR
#define SROW_NB_FAR 4             // far (L2) hops of one kind whose loads are in flight together
#endif
#ifndef SROW_NB_IN
#define SROW_NB_IN 1              // ... for sources inside the shared-memory tile (measured: 1 beats 2)
#endif
#define SROW_MAXG 64              // groups per chunk: 2^T, T <= 6
#define SMEM_LIMIT 232448        // 227 KB per CTA on sm_100

struct FastFactor {
  int W = 0, WT = 0, nvals = 0;
  int64_t n = 0;
  uint32_t *d_ell = nullptr;     // [WT][n] slot-major, padding entries point at column n (zero slot)
  double *d_vtab = nullptr;      // [nvals] magnitudes, vtab[0] = 0
  uint16_t *d_ell16 = nullptr;   // uniform factors with n < 32768: column | sign << 15 (halves the L2 index traffic);
                                 // [n][8] row-major when WT == 8 (one 16-byte load per row), else [WT][n]
  bool uniform = false;          // every stored magnitude identical -> y = v * sum(+-x)
  double vuni = 0.0;
  bool korder = false;           // single band: slots in the order of the row's eligible bath bits (UNI = 3 of k_fcol)
  double vk[EDGPU_MAX_SITES];    // V_k of this spin (korder)
};

struct SRowPlan {
  bool ok = false;
  int LR = 0, T = 0, nhigh = 0, nchunks = 0, cmax = 0, maxg = 0;
  SRowRec *d_recs = nullptr;     // group records, chunk by chunk
  int4 *d_chunks = nullptr;      // [nchunks] (first record, groups, local column begin, local column end)
  double *d_dr0 = nullptr, *d_dr1 = nullptr;   // direct mode: dfac_up[row] (+ Uloc when the up impurity is occupied)
  double vk[EDGPU_MAX_SITES];    // V_k of the dw spin, k = bath bit (1-based site k+1)
  double *d_vk = nullptr;
  size_t smem = 0;
  // sharded vector: dw hops the row kernel leaves out (source on another rank, or a low group cut by a boundary)
  bool lists = false;
  int nlist = 0, nzcols = 0;                   // entries; local columns that have entries
  int *d_lptr = nullptr;                       // [qdw+1] entry range of every local column
  int *d_lown = nullptr, *d_lcol = nullptr;    // entry: owner rank, column inside the owner's shard
  double *d_lamp = nullptr;
  unsigned char *d_lflag = nullptr;            // per local column: 1 = not written by k_srow (cut group: starts from the
                                               // diagonal in pass 2), 2 = has entries
  // halo (push model): the remote entries of rank q are numbered in list order = slots of q's halo buffers
  int nslot = 0, maxslot = 0;                  // slots of this rank; largest slot count of any rank (symmetric layout)
  int *d_lcol2 = nullptr;                      // entry: halo slot (remote owner) or local column (own)
  int npush = 0;                               // (peer, slot, local column) triples this rank stores every H*v
  int *d_pdst = nullptr, *d_pslot = nullptr, *d_psrc = nullptr;
  int nwin = 1;                                // column windows of the targets: triples sorted by (window, source column)
  int pwin[EDGPU_MAX_WINDOWS + 1] = {0};       // triple range of every window
  unsigned int *d_pushctr = nullptr;           // [window] CTAs of k_halo_push that are done (the last one raises the flags)
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
// pass 2: whole-column kernel
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
  // MODE == 2 (Lanczos epilogue): w = sx*(y + F x) - cprev*xp, written over xp; partials of (sx*x).w.  vect != nullptr
  // (second sweep of sp_lanc_eigh): also w -= sw_a*(sx*x) and vect += sw_zk*(sx*x), scalars in the LancState
  double *xp;
  double *vect;
  const LancState *st;
  double *partials;
  // LISTS (sharded vector): the dw hops the row pass left out, entries lptr[j] .. lptr[j+1] of local column j: amplitude
  // lamp, source column lcol of xb[lown] (a slot of the halo buffer for a remote owner, the local shard for an own
  // cut group).  lflag[j] & 1: the row pass did not write column j at all, so y(:, j) starts from the diagonal term
  // here (diagmode 1 stored, 2 recomputed).
  const int *lptr, *lown, *lcol;
  const double *lamp;
  const double *xb[EDGPU_MAXP];
  const unsigned char *lflag;
  int diagmode;
  // UNI == 3 (single band, level-dependent V_k): slot s of a row is its s-th eligible bath bit, amplitude V_k from the word
  int ns;
  double vk[EDGPU_MAX_SITES];
};

// diagonal of element (r, local column j) for the columns the row pass skipped (runtime form of DIAG 1 / 2)
__device__ __forceinline__ double fcol_init_diag(const FColArgs &a, int64_t j, int r) {
  if (a.diagmode == 1) return __ldcs(a.diag + j * (int64_t)a.n + r);
  double d = __ldg(a.dfac_c + r) + __ldg(a.dfac_s + a.coloff + j);
  const uint32_t mc = (uint32_t)__ldg(a.map_c + r), ms = (uint32_t)__ldg(a.map_s + a.coloff + j);
  for (int o = 0; o < a.norb; o++)
    if ((mc >> o) & 1u)
      for (int q = 0; q < a.norb; q++)
        if ((ms >> q) & 1u) d += (o == q) ? a.uloc[o] : a.ust;
  return d;
}
// the listed dw hops of one column (LISTS): a small table per pipeline stage in shared memory, written by the thread
// that issues the column's bulk copy (so it is published by the same mbarrier).  Four entries inline (the usual
// case: one per rank-crossing bath bit); the rest -- columns of a low group cut by a rank boundary list all their
// hops -- through the global tables.
struct ColEnt { const double *p; double a; };
struct ColTab {
  int ne, e0, flag, pad;
  ColEnt ent[4];
};
static_assert(sizeof(ColTab) == 80, "ColTab layout");
__device__ __forceinline__ void coltab_fill(ColTab &T, const FColArgs &a, int64_t j) {
  const int f = (int)__ldg(a.lflag + j);
  T.flag = f;
  T.e0 = __ldg(a.lptr + j);
  T.ne = (f & 2) ? __ldg(a.lptr + j + 1) - T.e0 : 0;
  for (int q = 0; q < 4 && q < T.ne; q++) {
    T.ent[q].p = a.xb[__ldg(a.lown + T.e0 + q)] + (size_t)__ldg(a.lcol + T.e0 + q) * a.n;
    T.ent[q].a = __ldg(a.lamp + T.e0 + q);
  }
}
__device__ __forceinline__ double coltab_at(const ColTab &T, const FColArgs &a, int ne, int r) {
  double z = 0.0;
#pragma unroll
  for (int q = 0; q < 4; q++)
    if (q < ne) { const ColEnt e = T.ent[q]; z = fma(e.a, __ldcs(e.p + r), z); }
  for (int e = T.e0 + 4; e < T.e0 + ne; e++)
    z = fma(__ldg(a.lamp + e), __ldcs(a.xb[__ldg(a.lown + e)] + (size_t)__ldg(a.lcol + e) * a.n + r), z);
  return z;
}
// MODE: 0 = y = F x, 1 = y += F x, 2 = y += F x fused with the first Lanczos vector update (y is only read)
// UNI: 0 = general (value table, 4-byte entries), 1 = uniform magnitude with 4-byte entries, 2 = uniform with 2-byte
// entries, 3 = single-band star geometry with level-dependent V_k: 2-byte entries in the order of the row's eligible
// bath bits (the ones whose occupation differs from the impurity's), so the amplitude of slot s is V_k of the s-th
// such bit of the row's word -- 20 bytes of index data per row instead of 32
template <int WT, int DIAG, int UNI, int MODE, bool LISTS>
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
  ColTab *tab = reinterpret_cast<ColTab *>(red + 32);            // [2] listed hops of the column in each stage (LISTS)
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
  if (UNI == 3 && tid < 34) vtab[tid] = (tid >= 2 && tid - 1 < a.ns) ? a.vk[tid - 1] : 0.0;   // vtab[j] = V_{j-1}; 0 = padding slot
  __syncthreads();
  const uint32_t colbytes = (uint32_t)n * 8u;
  const uint32_t bathmask = UNI == 3 ? (((1u << a.ns) - 1u) & ~1u) : 0u;
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
        if (LISTS) coltab_fill(tab[b], a, j);
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
      bool init = false;
      int ne = 0;
      const ColTab &T = tab[b];
      uint32_t en[WT];
      uint32_t mnext = 0;
      auto load_ell = [&](int row) {
        if (UNI == 3) mnext = (uint32_t)__ldg(a.map_c + row);
        if ((UNI == 2 || UNI == 3) && WT == 8) {                   // 8 two-byte entries = one LDG.128
          const uint4 q = __ldg(reinterpret_cast<const uint4 *>(a.ell16) + row);
          en[0] = q.x & 0xFFFFu; en[1] = q.x >> 16; en[2] = q.y & 0xFFFFu; en[3] = q.y >> 16;
          en[4 % WT] = q.z & 0xFFFFu; en[5 % WT] = q.z >> 16; en[6 % WT] = q.w & 0xFFFFu; en[7 % WT] = q.w >> 16;
        } else {
#pragma unroll
          for (int s = 0; s < WT; s++) en[s] = (UNI >= 2) ? (uint32_t)__ldg(a.ell16 + (size_t)s * n + row) : __ldg(a.ell + (size_t)s * n + row);
        }
      };
      if (tid < n) load_ell(tid);                                  // independent of the column: issued before the wait
      double ynext = 0.0, znext = 0.0;
      if (ACC && tid < n) ynext = __ldcs(yc + tid);
      mbar_wait(&bar[b], (uint32_t)((it >> 1) & 1));
      if (LISTS) {
        init = (T.flag & 1) != 0; ne = T.ne;
        if (ne && tid < n) znext = coltab_at(T, a, ne, tid);
      }
      for (int r = tid; r < n; r += NCT) {
        uint32_t e[WT];
#pragma unroll
        for (int s = 0; s < WT; s++) e[s] = en[s];
        const double yold = ynext, zold = znext;
        uint32_t E = 0;
        if (UNI == 3) E = (mnext & 1u) ? (~mnext & bathmask) : (mnext & bathmask);
        if (r + NCT < n) {                                         // software prefetch of the next row's inputs
          load_ell(r + NCT);
          if (ACC) ynext = __ldcs(yc + r + NCT);
          if (LISTS && ne) znext = coltab_at(T, a, ne, r + NCT);
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
          if (UNI == 3) {
            const int kk = __ffs((int)E);                          // 1-based bath bit of this slot, 0 once the row has run out
            E &= E - 1;
            acc = fma(flip_sign(vtab[kk], (e[s] & 0x8000u) << 16), xs[e[s] & 0x7FFFu], acc);
            continue;
          }
          const double xv = xs[e[s] & F_COL_MASK];
          if (UNI) acc += flip_sign(xv, e[s] & 0x80000000u);
          else acc = fma(flip_sign(vtab[(e[s] >> F_COL_BITS) & F_VID_MASK], e[s] & 0x80000000u), xv, acc);
        }
        if (UNI == 1 || UNI == 2) acc0 = fma(a.vuni, acc, acc0); else acc0 += acc;
        if (LISTS) {
          acc0 += init ? fcol_init_diag(a, j, r) * xs[r] : yold;
          acc0 += zold;
        } else if (ACC) {
          acc0 += yold;
        }
        if (MODE == 2) {
          double *wp = a.xp + j * (int64_t)n + r;
          double w = lsx * acc0 - lcp * __ldcs(wp);
          if (a.vect) {                                            // second sweep of sp_lanc_eigh: all scalars known
            const double xv = lsx * xs[r], lsa = __ldg(&a.st->sw_a), lzk = __ldg(&a.st->sw_zk);   // uniform, L1 resident
            w -= lsa * xv;
            double *vp = a.vect + j * (int64_t)n + r;
            __stcs(vp, fma(lzk, xv, __ldcs(vp)));
          }
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
// pass 2 for columns that do not fit one SM's shared memory (Ns = 18: 389 KB): a CLUSTER of two CTAs holds
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

template <int WT, int UNI, int MODE, bool LISTS>
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
  ColTab *tab = reinterpret_cast<ColTab *>(red + 32);            // listed hops of the current column (LISTS)
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
      if (LISTS) coltab_fill(tab[0], a, j);
      fence_proxy_async();
      mbar_expect_tx(&bar[0], bytes);
      const char *src = reinterpret_cast<const char *>(a.x + j * (int64_t)n + r0);
      for (uint32_t off = 0; off < bytes; off += 32768u) bulk_g2s(reinterpret_cast<char *>(buf) + off, src + off, min(32768u, bytes - off), &bar[0]);
      buf[nr] = 0.0;
    }
    mbar_wait(&bar[0], (uint32_t)(it & 1));
    cluster_sync_all();                                            // both halves of column j are in place
    double *yc = a.y + j * (int64_t)n + r0;
    bool init = false;
    int ne = 0;
    const ColTab &T = tab[0];
    if (LISTS) { init = (T.flag & 1) != 0; ne = T.ne; }
    for (int r = tid; r < nr; r += FCOL_THREADS) {
      uint32_t e[WT];
#pragma unroll
      for (int s = 0; s < WT; s++) e[s] = __ldg(a.ell + (size_t)s * n + r0 + r);
      double yold = 0.0, zold = 0.0;
      if (ACC) yold = __ldcs(yc + r);
      if (LISTS && ne) zold = coltab_at(T, a, ne, r0 + r);
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
      if (LISTS) {
        acc0 += init ? fcol_init_diag(a, j, r0 + r) * buf[r] : yold;
        acc0 += zold;
      } else if (ACC) {
        acc0 += yold;
      }
      if (MODE == 2) {
        double *wp = a.xp + j * (int64_t)n + r0 + r;
        double w = lsx * acc0 - lcp * __ldcs(wp);
        if (a.vect) {
          const double xv = lsx * buf[r], lsa = __ldg(&a.st->sw_a), lzk = __ldg(&a.st->sw_zk);
          w -= lsa * xv;
          double *vp = a.vect + j * (int64_t)n + r0 + r;
          __stcs(vp, fma(lzk, xv, __ldcs(vp)));
        }
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
// pass 1: structured row-tile kernel for the single-band star geometry
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
  const double *x;               // local shard, column 0
  double *y;
  int n;                         // DimUp (contiguous, even)
  uint32_t n8;                   // column stride in bytes
  int nchunks, tbits, nhigh, cpad;
  const int4 *chunks;
  const SRowRec *recs;
  double vk[EDGPU_MAX_SITES];    // vk[k], k = 1 .. Ns-1 (compile-time indices only: constant-bank operands)
  const double *vkd;             // the same table in global memory (run-time indices)
  const double *diag;            // DIAG == 1: local spH0d, one value per element
  const double *dr0, *dr1;       // DIAG == 2: per-row tables, dw impurity empty / occupied
  const double *dfac_s;          // DIAG == 2: per-column table of the WHOLE dw basis (padded by 2 doubles)
  int c0;                        // first global column of this shard
};

// what a consumer warp knows about the tile it works on
struct SRowTile {
  const double *tl;              // shared-memory tile + lane: local column cb + c at tl[c * 32]
  const double *dsc;             // shared: dfac_s of the tile's columns (DIAG == 2)
  const double *xrow;            // x + row of this lane (row clamped into the matrix)
  double *yrow;                  // y + row of this lane + first column of the tile
  const double *dgrow;           // diag + row + first column of the tile (DIAG == 1)
  const double *vhigh;           // shared: vhigh[kk] = V_{LR+kk}
  uint32_t n8, inmask, farmask;
  bool active;                   // this lane's row exists (stores only)
  double drow0, drow1;
};

// p + j columns (column stride n8 bytes); one IMAD.WIDE.U32
__device__ __forceinline__ const double *col_at(const double *p, uint32_t n8, uint32_t j) {
  return reinterpret_cast<const double *>(reinterpret_cast<const char *>(p) + (uint64_t)n8 * (uint64_t)j);
}
__device__ __forceinline__ double *col_at(double *p, uint32_t n8, uint32_t j) {
  return reinterpret_cast<double *>(reinterpret_cast<char *>(p) + (uint64_t)n8 * (uint64_t)j);
}

// All hops of ONE kind of a low group, NB at a time so that their loads are in flight together.  m: the hopped
// high bits.  BK = the bath bit is occupied in the target group: targets are the columns with the impurity
// empty, sources lo|1 in class N+1 of the partner group; otherwise targets have the impurity occupied, sources
// lo&~1 in class N-1.  SM = the partner group is in the shared-memory tile (column stride 32 doubles, immediate
// offsets), else in global memory (column stride n8 bytes).  pc[kk] = first column of the partner group (tile-local
// resp. shard-local), par bit kk = parity of the occupied high bits below kk.  A slot beyond the last hop gets a
// zero amplitude and re-reads one line of x that is in L1 anyway (zero column stride): no divergent control flow,
// no predicate set-up per load, no extra L2 traffic.
template <int LR, int N, bool BK, bool SM, int NB>
__device__ __forceinline__ void srow_hops(uint32_t m, const int32_t *pc, const double *vhigh, uint32_t par, const double *src0,
                                          uint32_t n8, double (&acc)[lowtab::binom(LR, N)]) {
  if constexpr ((BK && N == LR) || (!BK && N == 0)) {
    return;                                                        // no such targets in this class
  } else {
    constexpr int CNT = lowtab::binom(LR, N);
    constexpr int HB = BK ? lowtab::binom(LR - 1, N) : lowtab::binom(LR - 1, N - 1);
    while (m) {
      const double *p[NB];
      double amp[NB];
      uint32_t st[NB];
#pragma unroll
      for (int q = 0; q < NB; q++) {
        const bool on = m != 0;
        const int kk = on ? __ffs((int)m) - 1 : 0;
        m &= m - 1;
        const uint32_t col = on ? (uint32_t)pc[kk] : 0u;
        st[q] = on ? n8 : 0u;
        p[q] = SM ? src0 + col * SROW_R : col_at(src0, n8, col);
        const double v = vhigh[kk];
        amp[q] = on ? (((par >> kk) & 1u) ? -v : v) : 0.0;
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
            if constexpr (SM) v[q][hi] = p[q][j * SROW_R];
            else v[q][hi] = *col_at(p[q], st[q], j);
          }
        });
      });
      static_for<NB>([&](auto bc) {
        constexpr int q = decltype(bc)::value;
        static_for<CNT>([&](auto ic) {
          constexpr int i = decltype(ic)::value;
          constexpr int lo = lowtab::pat(LR, N, i);
          if constexpr (((lo & 1) == 0) == BK) {
            constexpr int par_lo = lowtab::popc(lo >> 1) & 1;
            constexpr int hi = lowtab::half_index(LR, N, i, BK ? 0 : 1);
            if constexpr (par_lo) acc[i] = fma(-amp[q], v[q][hi], acc[i]);
            else acc[i] = fma(amp[q], v[q][hi], acc[i]);
          }
        });
      });
    }
  }
}

// one low group of class N (N electrons in the LR low bits): record hd + partner columns pc (shared memory).
// vk: the hybridisations as kernel-parameter constants (vk[kb] with a compile-time kb is a constant-bank operand).
template <int LR, int N, int DIAG>
__device__ __forceinline__ void srow_group(const SRowTile &k, const int4 hd, const int32_t *pc, const double (&vk)[EDGPU_MAX_SITES]) {
  constexpr int CNT = lowtab::binom(LR, N);
  double acc[CNT];
  const int lb = hd.x;
  const uint32_t h = (uint32_t)hd.z & 0xFFFFu, ex = (uint32_t)hd.z >> 16, par = (uint32_t)hd.w;
  const double *tl = k.tl + lb * SROW_R;
  {
    // diagonal (direct mode: factorised tables staged in shared memory) and the hops among the low bits, register
    // to register; the group's own values are dead after this block
    double xv[CNT];
    static_for<CNT>([&](auto ic) { constexpr int i = decltype(ic)::value; xv[i] = tl[i * SROW_R]; });
    static_for<CNT>([&](auto ic) {
      constexpr int i = decltype(ic)::value;
      constexpr int lo = lowtab::pat(LR, N, i);
      if (DIAG == 2) acc[i] = (((lo & 1) ? k.drow1 : k.drow0) + k.dsc[lb + i]) * xv[i];
      else acc[i] = 0.0;
    });
    static_for<CNT>([&](auto ic) {
      constexpr int i = decltype(ic)::value;
      constexpr int lo = lowtab::pat(LR, N, i);
      static_for<LR - 1>([&](auto kc) {
        constexpr int kb = decltype(kc)::value + 1;
        if constexpr (((lo >> kb) & 1) != (lo & 1)) {
          constexpr int lo2 = lo ^ (1 | (1 << kb));
          constexpr int j = lowtab::rank(lo2);
          constexpr int par_lo = lowtab::popc(lo & ((1 << kb) - 2)) & 1;
          if constexpr (par_lo) acc[i] = fma(-vk[kb], xv[j], acc[i]);
          else acc[i] = fma(vk[kb], xv[j], acc[i]);
        }
      });
    });
  }
  // hops on the top bits: sources in L2
  srow_hops<LR, N, true, false, SROW_NB_FAR>(ex & h & k.farmask, pc, k.vhigh, par, k.xrow, k.n8, acc);
  srow_hops<LR, N, false, false, SROW_NB_FAR>(ex & ~h & k.farmask, pc, k.vhigh, par, k.xrow, k.n8, acc);
  // hops on the T lowest high bits: the partner group is in the tile
  srow_hops<LR, N, true, true, SROW_NB_IN>(ex & h & k.inmask, pc, k.vhigh, par, k.tl, k.n8, acc);
  srow_hops<LR, N, false, true, SROW_NB_IN>(ex & ~h & k.inmask, pc, k.vhigh, par, k.tl, k.n8, acc);
  if (DIAG == 1) {                                                 // streamed spH0d: one more pass over the group
    const double *dp = col_at(k.dgrow, k.n8, (uint32_t)lb);
    double dg[CNT];
    static_for<CNT>([&](auto ic) { constexpr int i = decltype(ic)::value; dg[i] = __ldcs(col_at(dp, k.n8, i)); });
    static_for<CNT>([&](auto ic) { constexpr int i = decltype(ic)::value; acc[i] = fma(dg[i], tl[i * SROW_R], acc[i]); });
  }
  if (k.active) {
    double *yp = col_at(k.yrow, k.n8, (uint32_t)lb);
    static_for<CNT>([&](auto ic) { constexpr int i = decltype(ic)::value; __stcs(col_at(yp, k.n8, i), acc[i]); });
  }
}

template <int LR, int DIAG, int... NS>
__device__ __forceinline__ void srow_dispatch(const SRowTile &k, const int4 hd, const int32_t *pc, const double (&vk)[EDGPU_MAX_SITES],
                                              std::integer_sequence<int, NS...>) {
  ((hd.y == NS ? (srow_group<LR, NS, DIAG>(k, hd, pc, vk), 0) : 0), ...);
}

// 15 consumer warps + 1 producer warp.  full[b]: the TMA copies of buffer b have landed; empty[b]: every
// consumer warp is done with buffer b.  Consumer warps take the low groups of the tile from a shared
// counter (largest first) and run ahead into the next buffer without a CTA-wide barrier.
template <int LR, int DIAG>
__global__ void __launch_bounds__(SROW_THREADS, 1) k_srow(const __grid_constant__ CUtensorMap tmx, SRowArgs a) {
  extern __shared__ __align__(128) unsigned char smraw[];
  constexpr int S = SROW_STAGES;
  const int tsz = a.cpad * SROW_R;                                // doubles per tile buffer
  double *tile0 = reinterpret_cast<double *>(smraw);              // [S][cpad * 32]
  double *dsc0 = tile0 + (size_t)S * tsz;                         // [S][cpad + 2]
  double *drw0 = dsc0 + S * (a.cpad + 2);                         // [S][2][32]
  double *vhigh = drw0 + S * 2 * SROW_R;                          // [32]
  SRowRec *rec0 = reinterpret_cast<SRowRec *>(vhigh + 32);        // [S][SROW_MAXG], 16-byte aligned
  uint64_t *bar = reinterpret_cast<uint64_t *>(rec0 + S * SROW_MAXG);   // full[S], empty[S]
  int *gctr = reinterpret_cast<int *>(bar + 2 * S);               // [S]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int b = 0; b < S; b++) {
      mbar_init(&bar[b], 1);
      mbar_init(&bar[S + b], SROW_CONSUMERS / 32);
      gctr[b] = 0;
    }
    fence_barrier_init();
  }
  if (tid < 32) vhigh[tid] = (tid < a.nhigh) ? __ldg(a.vkd + LR + tid) : 0.0;
  __syncthreads();

  const int nrb = (a.n + SROW_R - 1) / SROW_R;
  const int nitems = nrb * a.nchunks;                             // < 2^31: at most 2^19 items at Ns = 18
  const int G = (int)gridDim.x;
  if (warp == SROW_CONSUMERS / 32) {
    // ---- producer warp ----
    for (int it = 0;; it++) {
      const int t = (int)blockIdx.x + it * G;
      if (t >= nitems) break;
      const int b = it % S;
      const int use = it / S;
      if (use >= 1) mbar_wait(&bar[S + b], (uint32_t)((use - 1) & 1));
      const int rb = t / a.nchunks;
      const int4 ch = __ldg(a.chunks + (t % a.nchunks));
      const int nc = ch.w - ch.z;
      const int nops = (nc + SROW_BC - 1) / SROW_BC;
      const int i0 = rb * SROW_R;
      const int nr = min(SROW_R, a.n - i0);
      const int lead = (a.c0 + ch.z) & 1;                          // bulk copies start on a 16-byte boundary
      const uint32_t dsc_bytes = (uint32_t)((lead + nc + 1) & ~1) * 8u;
      const uint32_t rec_bytes = (uint32_t)ch.y * (uint32_t)sizeof(SRowRec);
      if (lane == 0) {
        gctr[b] = 0;
        fence_proxy_async();
        uint32_t bytes = (uint32_t)nops * (uint32_t)(SROW_BC * SROW_R * 8) + rec_bytes;
        if (DIAG == 2) bytes += dsc_bytes + 2u * (uint32_t)nr * 8u;
        mbar_expect_tx(&bar[b], bytes);
      }
      __syncwarp();
      if (lane < nops)
        tma_load_2d(tile0 + (size_t)b * tsz + (size_t)lane * SROW_BC * SROW_R, &tmx, i0, ch.z + lane * SROW_BC, &bar[b]);
      if (lane == 28) bulk_g2s(rec0 + b * SROW_MAXG, a.recs + ch.x, rec_bytes, &bar[b]);
      if (DIAG == 2) {
        if (lane == 29) bulk_g2s(dsc0 + (size_t)b * (a.cpad + 2), a.dfac_s + (a.c0 + ch.z - lead), dsc_bytes, &bar[b]);
        if (lane == 30) bulk_g2s(drw0 + (size_t)b * 2 * SROW_R, a.dr0 + i0, (uint32_t)nr * 8u, &bar[b]);
        if (lane == 31) bulk_g2s(drw0 + (size_t)b * 2 * SROW_R + SROW_R, a.dr1 + i0, (uint32_t)nr * 8u, &bar[b]);
      }
    }
    return;
  }
  // ---- consumer warps ----
  SRowTile k;
  k.vhigh = vhigh;
  k.n8 = a.n8;
  k.inmask = (1u << a.tbits) - 1u;
  k.farmask = ((1u << a.nhigh) - 1u) & ~k.inmask;
  for (int it = 0;; it++) {
    const int t = (int)blockIdx.x + it * G;
    if (t >= nitems) break;
    const int b = it % S;
    const int rb = t / a.nchunks;
    const int4 ch = __ldg(a.chunks + (t % a.nchunks));
    const int i0 = rb * SROW_R;
    const int row = min(i0 + lane, a.n - 1);                         // clamp: loads of a ragged last block stay in range
    k.tl = tile0 + (size_t)b * tsz + lane;
    k.dsc = dsc0 + (size_t)b * (a.cpad + 2) + ((a.c0 + ch.z) & 1);
    k.xrow = a.x + row;
    k.yrow = col_at(a.y + row, a.n8, (uint32_t)ch.z);
    k.dgrow = col_at(a.diag + row, a.n8, (uint32_t)ch.z);
    k.active = (i0 + lane) < a.n;
    const SRowRec *recs = rec0 + b * SROW_MAXG;
    mbar_wait(&bar[b], (uint32_t)((it / S) & 1));
    k.drow0 = 0.0; k.drow1 = 0.0;
    if (DIAG == 2) {
      k.drow0 = drw0[b * 2 * SROW_R + lane];
      k.drow1 = drw0[b * 2 * SROW_R + SROW_R + lane];
    }
    for (;;) {
      int g = 0;
      if (lane == 0) g = atomicAdd(&gctr[b], 1);
      g = __shfl_sync(0xffffffffu, g, 0);
      if (g >= ch.y) break;
      const int4 hd = *reinterpret_cast<const int4 *>(&recs[g]);
      srow_dispatch<LR, DIAG>(k, hd, recs[g].pc, a.vk, std::make_integer_sequence<int, LR + 1>{});
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar[S + b]);                       // this warp is done with buffer b
  }
}

// ---------------------------------------------------------------------------------------------
// halo, producer side: this rank's source columns go to the halo buffers of the ranks that list them, moved by the
// TMA engine both ways -- cp.async.bulk global -> shared (local read) and shared -> global (posted writes to the
// peer-mapped address over NVLink), a ring of 16 KB stages driven by ONE thread per CTA (measured against 16-byte
// register stores: tools/nvlink_push.cu, 667 vs 523 GB/s with 32 CTAs).  Work item = (triple, 16 KB piece); triples
// are sorted by source column, so a column that several peers need is read from HBM once.  The last CTA to finish raises this
// rank's arrival flag (= the H*v epoch) on every peer: fence.sys by every CTA before it counts itself, fence.sys by
// the last one before the flag stores (cumulativity makes all the data visible before the flag).
// ---------------------------------------------------------------------------------------------
struct PushArgs {
  const double *x;               // local shard
  int n, npush;
  const int *pdst, *pslot, *psrc;
  double *hb[EDGPU_MAXP];        // rank p's halo buffer of this epoch
  unsigned long long *flag[EDGPU_MAXP];   // rank p's arrival flag of THIS rank, nullptr = no flag (own rank, emulation)
  unsigned long long epoch;
  unsigned int *ctr;
};
#define PUSH_SB 16384             // bytes per stage
#define PUSH_STAGES 4
#define PUSH_SMEM (PUSH_STAGES * PUSH_SB + 64)
__global__ void __launch_bounds__(32) k_halo_push(PushArgs a) {
  extern __shared__ __align__(128) unsigned char psm[];
  uint64_t *bar = reinterpret_cast<uint64_t *>(psm + (size_t)PUSH_STAGES * PUSH_SB);
  if (threadIdx.x != 0) return;                                    // the TMA engine does the work: one thread drives it
  for (int s = 0; s < PUSH_STAGES; s++) mbar_init(&bar[s], 1);
  fence_barrier_init();
  const int64_t colbytes = (int64_t)a.n * 8;
  const int npc = (int)((colbytes + PUSH_SB - 1) / PUSH_SB);
  const int64_t nitems = (int64_t)a.npush * npc, G = gridDim.x;
  const int64_t my = nitems > blockIdx.x ? (nitems - blockIdx.x + G - 1) / G : 0;
  auto piece = [&](int64_t k, int *e, int64_t *off, uint32_t *bytes) {
    const int64_t it = blockIdx.x + k * G;
    *e = (int)(it / npc);
    *off = (it % npc) * PUSH_SB;
    *bytes = (uint32_t)min((int64_t)PUSH_SB, colbytes - *off);
  };
  auto load = [&](int64_t k) {
    int e; int64_t off; uint32_t bytes;
    piece(k, &e, &off, &bytes);
    const int s = (int)(k % PUSH_STAGES);
    mbar_expect_tx(&bar[s], bytes);
    bulk_g2s(psm + (size_t)s * PUSH_SB, reinterpret_cast<const char *>(a.x + (size_t)__ldg(a.psrc + e) * a.n) + off, bytes, &bar[s]);
  };
  int64_t issued = 0;
  for (; issued < my && issued < PUSH_STAGES; issued++) load(issued);
  for (int64_t k = 0; k < my; k++) {
    const int s = (int)(k % PUSH_STAGES);
    mbar_wait(&bar[s], (uint32_t)((k / PUSH_STAGES) & 1));
    int e; int64_t off; uint32_t bytes;
    piece(k, &e, &off, &bytes);
    char *dst = reinterpret_cast<char *>(a.hb[__ldg(a.pdst + e)] + (size_t)__ldg(a.pslot + e) * a.n) + off;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(psm + (size_t)s * PUSH_SB)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    if (issued < my) {                                             // the next load reuses this stage: the store must have read it
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      load(issued);
      issued++;
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");       // every store of this CTA is complete ...
  asm volatile("fence.proxy.async;" ::: "memory");                // ... and ordered before the generic-proxy flag protocol
  __threadfence_system();
  if (atomicAdd(a.ctr, 1u) == gridDim.x - 1) {
    *a.ctr = 0u;                                                   // ready for the next launch (stream ordered)
    __threadfence_system();
    for (int p = 0; p < EDGPU_MAXP; p++)
      if (a.flag[p]) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.flag[p]), "l"(a.epoch) : "memory");
  }
}
// consumer side: wait until every peer's columns of this epoch have arrived (flags live in local memory)
__global__ void k_halo_wait(const unsigned long long *flags, int nranks, int me, unsigned long long epoch) {
  const int p = threadIdx.x;
  if (p < nranks && p != me) {
    const long long t0 = clock64();
    for (;;) {
      unsigned long long v;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + p) : "memory");
      if (v >= epoch) break;
      if (clock64() - t0 > 1200000000000LL) __trap();              // ~10 min (ranks may arrive late after host work): a peer died -- fail loudly instead of hanging
      __nanosleep(200);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------
static int pack_fast(edgpu_ctx *c, const Factor &f, FastFactor &ff) {
  std::vector<int32_t> rp((size_t)f.n + 1), cols((size_t)std::max<int64_t>(f.nnz, 1)), map((size_t)f.n);
  std::vector<double> vals((size_t)std::max<int64_t>(f.nnz, 1));
  CK(cudaMemcpy(rp.data(), f.d_rowptr, rp.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(map.data(), f.d_map, map.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
  if (f.nnz) {
    CK(cudaMemcpy(cols.data(), f.d_cols, (size_t)f.nnz * sizeof(int32_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(vals.data(), f.d_vals, (size_t)f.nnz * sizeof(double), cudaMemcpyDeviceToHost));
  }
  ff.n = f.n;
  // Single band: every entry (row <- col) differs from the row's word in the impurity bit and ONE bath bit k and
  // carries +-V_k.  Then the slots of a row can follow its eligible bath bits (those whose occupation differs from
  // the impurity's) in ascending order, and no value id has to be stored: slot s <-> s-th eligible bit.
  const int ns = c->ns, spin = (&f == &c->dw) ? 1 : 0;
  for (int k = 0; k < EDGPU_MAX_SITES; k++) ff.vk[k] = 0.0;
  for (int k = 1; k < ns; k++) ff.vk[k] = spin ? c->dp.bv_dw[k - 1] : c->dp.bv_up[k - 1];
  const uint32_t bath = ((1u << ns) - 1u) & ~1u;
  bool star = c->dp.norb == 1;
  int wstar = 1;
  for (int64_t i = 0; i < f.n && star; i++) {
    const uint32_t m = (uint32_t)map[(size_t)i];
    wstar = std::max(wstar, __builtin_popcount((m & 1u) ? (~m & bath) : (m & bath)));
    for (int32_t p = rp[(size_t)i]; p < rp[(size_t)i + 1]; p++) {
      const uint32_t x = m ^ (uint32_t)map[(size_t)cols[(size_t)p]];
      if (!(x & 1u) || __builtin_popcount(x) != 2) { star = false; break; }
      const int k = __builtin_ctz(x & ~1u);
      if (fabs(vals[(size_t)p]) != fabs(ff.vk[k])) { star = false; break; }
    }
  }
  ff.korder = star && wstar <= 16;
  ff.W = ff.korder ? wstar : std::max(f.maxrow, 1);
  ff.WT = ff.W <= 8 ? 8 : (ff.W <= 12 ? 12 : 16);
  if (ff.W > 16) return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "factor row too long for the fast column kernel");
  std::map<double, int> ids;
  std::vector<double> vtab(1, 0.0);
  std::vector<uint32_t> ell((size_t)ff.WT * f.n, (uint32_t)f.n);       // padding: zero slot, value id 0
  for (int64_t i = 0; i < f.n; i++) {
    const uint32_t m = (uint32_t)map[(size_t)i];
    uint32_t E = (m & 1u) ? (~m & bath) : (m & bath);
    int k = 0;
    auto put = [&](int slot, int32_t p) {
      const double av = vals[(size_t)p] < 0 ? -vals[(size_t)p] : vals[(size_t)p];
      auto it = ids.find(av);
      int id;
      if (it == ids.end()) { id = (int)vtab.size(); ids[av] = id; vtab.push_back(av); }
      else id = it->second;
      if (id >= F_MAXVALS) return false;
      ell[(size_t)slot * f.n + i] = (uint32_t)cols[(size_t)p] | ((uint32_t)id << F_COL_BITS) | (vals[(size_t)p] < 0 ? 0x80000000u : 0u);
      return true;
    };
    if (ff.korder) {
      for (; E; E &= E - 1, k++) {                                  // slot k = k-th eligible bath bit; V_k = 0: padding entry
        const uint32_t m2 = m ^ 1u ^ (1u << __builtin_ctz(E));
        for (int32_t p = rp[(size_t)i]; p < rp[(size_t)i + 1]; p++)
          if ((uint32_t)map[(size_t)cols[(size_t)p]] == m2 && !put(k, p))
            return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "too many distinct matrix elements for the fast column kernel");
      }
    } else {
      for (int32_t p = rp[(size_t)i]; p < rp[(size_t)i + 1]; p++, k++)
        if (!put(k, p)) return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "too many distinct matrix elements for the fast column kernel");
    }
  }
  ff.nvals = (int)vtab.size();
  ff.uniform = (ff.nvals == 2);
  ff.vuni = ff.uniform ? vtab[1] : 0.0;
  CK(cudaMalloc(&ff.d_ell, ell.size() * sizeof(uint32_t)));
  CK(cudaMalloc(&ff.d_vtab, vtab.size() * sizeof(double)));
  CK(cudaMemcpy(ff.d_ell, ell.data(), ell.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ff.d_vtab, vtab.data(), vtab.size() * sizeof(double), cudaMemcpyHostToDevice));
  if ((ff.uniform || ff.korder) && f.n < 32768) {
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

static size_t fcol_smem(int64_t n) { return (size_t)2 * (n + 2) * 8 + F_MAXVALS * 8 + 32 + 32 * 8 + 2 * sizeof(ColTab); }
static size_t fcol2_smem(int64_t n) { return (size_t)((((n >> 1) + 1) & ~(int64_t)1) + 2) * 8 + F_MAXVALS * 8 + 16 + 32 * 8 + sizeof(ColTab); }
static size_t srow_smem(int cpad) {
  size_t b = (size_t)SROW_STAGES * ((size_t)cpad * SROW_R * 8 + (size_t)(cpad + 2) * 8 + 2 * SROW_R * 8 + SROW_MAXG * sizeof(SRowRec) + 16 + 4) +
             32 * 8 + 64;
  return (b + 15) & ~(size_t)15;
}

// ---- pure host arithmetic of the row-kernel plan (also reachable without a GPU through
// edgpu_selftest_srow_plan, so that the CPU-only tests can check the chunk / record / list logic) ----------
//
// dw word = (h << LR) | lo.  Group = all words with the same h (class N = ndw - popc(h) electrons in the low bits),
// chunk = all groups with the same h >> T.  The basis is sorted by word, so groups and chunks are runs of columns.
// On rank `rank` of `nranks` a group is "whole" when all its columns are local; the kernel works on whole groups only.
int srow_plan_host(int ns, int ndw, int64_t dimdw, int nranks, int rank, int lr, int tbits_opt, SRowHostPlan &hp) {
  hp = SRowHostPlan();
  const int LR = (lr == 4 || lr == 5) ? lr : 5;
  if (ns <= LR || ns - LR > 16 || nranks > EDGPU_MAXP || dimdw >= (1 << 30)) return 0;
  hp.LR = LR;
  hp.nhigh = ns - LR;
  const int P = nranks;
  hp.coloffs.assign((size_t)P + 1, 0);
  for (int p = 0; p <= P; p++) {
    int64_t q = 0, off = dimdw;
    if (p < P) edgpu_split(dimdw, P, p, &q, &off);
    hp.coloffs[(size_t)p] = (int)off;
  }
  const int c0 = hp.coloffs[(size_t)rank], c1 = hp.coloffs[(size_t)rank + 1];
  const int nh = 1 << hp.nhigh;
  hp.jhi.assign((size_t)nh, -1);
  hp.gwhole.assign((size_t)nh, 0);
  int64_t col = 0;
  for (int h = 0; h < nh; h++) {
    const int nlow = ndw - __builtin_popcount((unsigned)h);
    if (nlow < 0 || nlow > LR) continue;
    const int sz = lowtab::binom(LR, nlow);
    hp.jhi[(size_t)h] = (int32_t)col;
    hp.gwhole[(size_t)h] = (col >= c0 && col + sz <= c1) ? 1 : 0;
    col += sz;
  }
  if (col != dimdw) return -1;                                     // the Lin table must cover the basis exactly
  // chunk width: the largest T whose tiles (two buffers of 32 rows) fit the shared memory, at most 2^T = SROW_MAXG groups
  int T = 0;
  for (int t = 0; t <= hp.nhigh && (1 << t) <= SROW_MAXG; t++) {
    const int cmax = lowtab::binom(LR + t, (LR + t) / 2);
    const int cpad = (cmax + SROW_BC - 1) / SROW_BC * SROW_BC;
    if (srow_smem(cpad) <= SMEM_LIMIT && t <= 5) T = t;            // T = 6 needs 16-row tiles: not built
  }
  if (tbits_opt > 0 && tbits_opt - 1 <= T) T = tbits_opt - 1;      // option value t+1 forces T = t (tests: small chunks)
  hp.T = T;
  const int nchunk_ids = 1 << (hp.nhigh - T);
  hp.cmax = 0; hp.maxg = 0;
  for (int ci = 0; ci < nchunk_ids; ci++) {
    // whole local groups of this chunk: a run of consecutive h (local columns are one interval)
    std::vector<int> hs;
    for (int u = 0; u < (1 << T); u++) {
      const int h = (ci << T) | u;
      if (hp.jhi[(size_t)h] >= 0 && hp.gwhole[(size_t)h]) hs.push_back(h);
    }
    if (hs.empty()) continue;
    auto gsize = [&](int h) { return lowtab::binom(LR, ndw - __builtin_popcount((unsigned)h)); };
    const int cb = hp.jhi[(size_t)hs.front()], ce = hp.jhi[(size_t)hs.back()] + gsize(hs.back());
    std::stable_sort(hs.begin(), hs.end(), [&](int a, int b) { return gsize(a) > gsize(b); });   // largest first
    const int rec_begin = (int)hp.recs.size();
    for (int h : hs) {
      SRowRec r;
      r.lb = hp.jhi[(size_t)h] - cb;
      r.N = ndw - __builtin_popcount((unsigned)h);
      uint32_t ex = 0, par = 0;
      for (int kk = 0; kk < 16; kk++) r.pc[kk] = 0;
      for (int kk = 0; kk < hp.nhigh; kk++) {
        const int h2 = h ^ (1 << kk);
        if (__builtin_popcount((unsigned)(h & ((1 << kk) - 1))) & 1) par |= 1u << kk;
        if (hp.jhi[(size_t)h2] < 0 || !hp.gwhole[(size_t)h2]) continue;   // no such group, or not (wholly) on this rank
        ex |= 1u << kk;
        r.pc[kk] = kk < T ? hp.jhi[(size_t)h2] - cb : hp.jhi[(size_t)h2] - c0;
      }
      r.hx = (uint32_t)h | (ex << 16);
      r.par = par;
      hp.recs.push_back(r);
    }
    hp.chunks.push_back(make_int4(rec_begin, (int)hs.size(), cb - c0, ce - c0));
    hp.cmax = std::max(hp.cmax, ce - cb);
    hp.maxg = std::max(hp.maxg, (int)hs.size());
  }
  hp.cpad = std::max(SROW_BC, (hp.cmax + SROW_BC - 1) / SROW_BC * SROW_BC);
  hp.smem = srow_smem(hp.cpad);
  if (hp.smem > SMEM_LIMIT || hp.maxg > SROW_MAXG) return 0;
  hp.ok = true;
  return 1;
}

// Source lists: every dw hop (target <- source) of a LOCAL target column that the row kernel does not apply, i.e. the
// target's or the source's low group is not whole on this rank, from the reference-order CSR of spH0dws
// (stored/H_dw.f90).  Columns of groups that are not whole are not written by the row kernel at all (lflag & 1).
// map: Hs(2)%map.
void srow_lists_host(SRowHostPlan &hp, int rank, int64_t dimdw, const int32_t *map, const int32_t *rp, const int32_t *cc, const double *vv) {
  const int P = (int)hp.coloffs.size() - 1;
  const int c0 = hp.coloffs[(size_t)rank], c1 = hp.coloffs[(size_t)rank + 1];
  hp.lptr.assign(1, 0); hp.lown.clear(); hp.lcol.clear(); hp.lamp.clear(); hp.lflag.clear(); hp.zcols.clear();
  if (P <= 1) return;
  (void)dimdw;
  auto owner_of = [&](int col) { int p = 0; while (p + 1 < P && col >= hp.coloffs[(size_t)p + 1]) p++; return p; };
  for (int t = c0; t < c1; t++) {
    const bool tw = hp.gwhole[(size_t)((uint32_t)map[t] >> hp.LR)] != 0;
    const size_t before = hp.lown.size();
    for (int32_t q = rp[(size_t)t]; q < rp[(size_t)t + 1]; q++) {
      const int s = cc[(size_t)q];
      if (tw && hp.gwhole[(size_t)((uint32_t)map[s] >> hp.LR)]) continue;      // k_srow applies it
      const int own = owner_of(s);
      hp.lown.push_back(own); hp.lcol.push_back(s - hp.coloffs[(size_t)own]); hp.lamp.push_back(vv[(size_t)q]);
    }
    hp.lptr.push_back((int)hp.lown.size());
    const bool has = hp.lown.size() > before;
    hp.lflag.push_back((unsigned char)((tw ? 0 : 1) | (has ? 2 : 0)));
    if (has) hp.zcols.push_back(t - c0);
  }
}

template <typename T>
static int to_device(T **d, const std::vector<T> &h) {
  const size_t n = std::max<size_t>(h.size(), 1);
  CK(cudaMalloc(d, n * sizeof(T)));
  if (!h.empty()) CK(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return EDGPU_OK;
}

// Halo tables of the push model.  Rank q's remote entries (owner != q), in list order, are the slots of q's halo
// buffers.  lcol2 (this rank's consumer view): slot for a remote entry, local column for an own one.  Triples (q, slot,
// local column) for every entry of another rank q that this rank owns, sorted by (column window of the target on q,
// source column): window w of rank q = its local columns [qdw*w/K, qdw*(w+1)/K), so that q's column pass of window w
// can start when the triples of window w have arrived.  hp_me: this rank's plan with its lists built.
int halo_tables_host(int ns, int ndw, int64_t dimdw, int P, int me, int lr, int tbits_opt, const SRowHostPlan &hp_me,
                     const int32_t *map, const int32_t *rp, const int32_t *cc, const double *vv, int K, std::vector<int> &lcol2,
                     std::vector<int> &pdst, std::vector<int> &pslot, std::vector<int> &psrc, int *pwin, int *nslot, int *maxslot) {
  lcol2.clear(); pdst.clear(); pslot.clear(); psrc.clear();
  *nslot = 0; *maxslot = 0;
  struct Tr { int win, src, dst, slot; };
  std::vector<Tr> tr;
  for (int q = 0; q < P; q++) {
    SRowHostPlan hq;
    const SRowHostPlan *h = &hp_me;
    if (q != me) {
      if (srow_plan_host(ns, ndw, dimdw, P, q, lr, tbits_opt, hq) != 1)
        return edgpu_set_err(EDGPU_ERR_INVALID, "internal: row-kernel plan of rank %d differs from rank %d's", q, me);
      srow_lists_host(hq, q, dimdw, map, rp, cc, vv);
      h = &hq;
    }
    const int64_t qq = (int64_t)h->lptr.size() - 1;                 // local columns of rank q
    int slot = 0;
    for (int64_t t = 0; t < qq; t++) {
      int w = 0;
      while (w + 1 < K && t >= qq * (w + 1) / K) w++;
      for (int e = h->lptr[(size_t)t]; e < h->lptr[(size_t)t + 1]; e++) {
        const bool remote = h->lown[(size_t)e] != q;
        if (q == me) lcol2.push_back(remote ? slot : h->lcol[(size_t)e]);
        else if (h->lown[(size_t)e] == me) tr.push_back(Tr{w, h->lcol[(size_t)e], q, slot});
        if (remote) slot++;
      }
    }
    if (q == me) *nslot = slot;
    *maxslot = std::max(*maxslot, slot);
  }
  std::stable_sort(tr.begin(), tr.end(), [](const Tr &a, const Tr &b) { return a.win != b.win ? a.win < b.win : a.src < b.src; });
  for (int w = 0; w <= K; w++) pwin[w] = 0;
  for (const Tr &t : tr) { pdst.push_back(t.dst); pslot.push_back(t.slot); psrc.push_back(t.src); pwin[t.win + 1]++; }
  for (int w = 0; w < K; w++) pwin[w + 1] += pwin[w];
  return EDGPU_OK;
}

static int build_srow(edgpu_ctx *c, SRowPlan &sr) {
  // single band, star geometry, no inter-orbital terms: every dw hop is bit 0 <-> bit k
  sr.ok = false;
  if (c->dp.norb != 1 || c->dp.jhflag || (c->dimup & 1)) return EDGPU_OK;
  SRowHostPlan hp;
  const int prc = srow_plan_host(c->ns, c->ndw, c->dimdw, c->nranks, c->rank, (int)c->opt_srow_lr, (int)c->opt_srow_t, hp);
  if (prc < 0) return edgpu_set_err(EDGPU_ERR_INVALID, "internal: Lin table does not cover the dw basis");
  if (prc == 0) return EDGPU_OK;
  sr.LR = hp.LR; sr.T = hp.T; sr.nhigh = hp.nhigh; sr.nchunks = (int)hp.chunks.size();
  sr.cmax = hp.cpad; sr.maxg = hp.maxg; sr.smem = hp.smem;
  for (int k = 0; k < EDGPU_MAX_SITES; k++) sr.vk[k] = 0.0;
  for (int k = 1; k < c->ns; k++) sr.vk[k] = c->dp.bv_dw[k - 1];
  TRY(to_device(&sr.d_vk, std::vector<double>(sr.vk, sr.vk + EDGPU_MAX_SITES)));
  TRY(to_device(&sr.d_recs, hp.recs));
  TRY(to_device(&sr.d_chunks, hp.chunks));
  if (c->up.d_dfac) {                                              // direct mode: per-row diagonal tables
    std::vector<double> d0((size_t)c->dimup + 32, 0.0), d1((size_t)c->dimup + 32, 0.0);
    std::vector<int32_t> mu((size_t)c->dimup);
    CK(cudaMemcpy(d0.data(), c->up.d_dfac, (size_t)c->dimup * sizeof(double), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(mu.data(), c->up.d_map, (size_t)c->dimup * sizeof(int32_t), cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < c->dimup; i++) d1[(size_t)i] = d0[(size_t)i] + ((mu[(size_t)i] & 1) ? c->dp.uloc[0] : 0.0);
    TRY(to_device(&sr.d_dr0, d0));
    TRY(to_device(&sr.d_dr1, d1));
  }
  sr.lists = false; sr.nlist = 0; sr.nzcols = 0;
  if (c->nranks > 1) {
    std::vector<int32_t> map((size_t)c->dw.n), rp((size_t)c->dw.n + 1), cc((size_t)std::max<int64_t>(c->dw.nnz, 1));
    std::vector<double> vv((size_t)std::max<int64_t>(c->dw.nnz, 1));
    CK(cudaMemcpy(map.data(), c->dw.d_map, map.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(rp.data(), c->dw.d_rowptr, rp.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (c->dw.nnz) {
      CK(cudaMemcpy(cc.data(), c->dw.d_cols, (size_t)c->dw.nnz * sizeof(int32_t), cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(vv.data(), c->dw.d_vals, (size_t)c->dw.nnz * sizeof(double), cudaMemcpyDeviceToHost));
    }
    srow_lists_host(hp, c->rank, c->dimdw, map.data(), rp.data(), cc.data(), vv.data());
    sr.lists = true;
    sr.nlist = (int)hp.lown.size();
    sr.nzcols = (int)hp.zcols.size();
    // halo slots and push triples: the plan and the lists are pure functions of (geometry, rank), so every rank
    // derives every other rank's lists itself -- nothing is exchanged
    std::vector<int> lcol2, pdst, pslot, psrc;
    // windows pay when the halo outlasts the row pass: from 4 ranks on (measured)
    sr.nwin = (int)std::max<int64_t>(1, std::min<int64_t>(c->opt_halo_windows > 0 ? c->opt_halo_windows : (c->nranks >= 4 ? 4 : 1), EDGPU_MAX_WINDOWS));
    TRY(halo_tables_host(c->ns, c->ndw, c->dimdw, c->nranks, c->rank, (int)c->opt_srow_lr, (int)c->opt_srow_t, hp, map.data(), rp.data(),
                         cc.data(), vv.data(), sr.nwin, lcol2, pdst, pslot, psrc, sr.pwin, &sr.nslot, &sr.maxslot));
    sr.npush = (int)pdst.size();
    TRY(to_device(&sr.d_lptr, hp.lptr));
    TRY(to_device(&sr.d_lown, hp.lown));
    TRY(to_device(&sr.d_lcol, hp.lcol));
    TRY(to_device(&sr.d_lcol2, lcol2));
    TRY(to_device(&sr.d_lamp, hp.lamp));
    TRY(to_device(&sr.d_lflag, hp.lflag));
    TRY(to_device(&sr.d_pdst, pdst));
    TRY(to_device(&sr.d_pslot, pslot));
    TRY(to_device(&sr.d_psrc, psrc));
    CK(cudaMalloc(&sr.d_pushctr, EDGPU_MAX_WINDOWS * sizeof(unsigned int)));
    CK(cudaMemset(sr.d_pushctr, 0, EDGPU_MAX_WINDOWS * sizeof(unsigned int)));
  }
  sr.ok = true;
  return EDGPU_OK;
}

int fast_plan_free(edgpu_ctx *c) {
  if (!c->fplan) return EDGPU_OK;
  for (int k = 0; k < 2; k++) { cudaFree(c->fplan->ff[k].d_ell); cudaFree(c->fplan->ff[k].d_ell16); cudaFree(c->fplan->ff[k].d_vtab); }
  SRowPlan &r = c->fplan->sr;
  cudaFree(r.d_recs); cudaFree(r.d_chunks); cudaFree(r.d_vk); cudaFree(r.d_dr0); cudaFree(r.d_dr1);
  cudaFree(r.d_lptr); cudaFree(r.d_lown); cudaFree(r.d_lcol); cudaFree(r.d_lamp); cudaFree(r.d_lflag);
  cudaFree(r.d_lcol2); cudaFree(r.d_pdst); cudaFree(r.d_pslot); cudaFree(r.d_psrc); cudaFree(r.d_pushctr);
  delete c->fplan;
  c->fplan = nullptr;
  return EDGPU_OK;
}

// the opt-in to > 48 KB of dynamic shared memory is a per-device function attribute: set once per context
template <int WT, int DIAG>
static cudaError_t set_fcol_attr() {
  cudaError_t e = cudaSuccess;
#define SETF(U, A, L) if (e == cudaSuccess) e = cudaFuncSetAttribute(k_fcol<WT, DIAG, U, A, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT)
  SETF(0, 0, false); SETF(1, 0, false); SETF(2, 0, false); SETF(3, 0, false);
  SETF(0, 1, false); SETF(1, 1, false); SETF(2, 1, false); SETF(3, 1, false);
  if (DIAG == 0) {
    SETF(0, 2, false); SETF(1, 2, false); SETF(2, 2, false); SETF(3, 2, false);
    SETF(0, 1, true); SETF(1, 1, true); SETF(2, 1, true); SETF(3, 1, true);
    SETF(0, 2, true); SETF(1, 2, true); SETF(2, 2, true); SETF(3, 2, true);
  }
#undef SETF
  return e;
}
template <int WT>
static cudaError_t set_fcol2_attr() {
  cudaError_t e = cudaSuccess;
#define SET2(U, M, L) if (e == cudaSuccess) e = cudaFuncSetAttribute(k_fcol2<WT, U, M, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT)
  SET2(0, 0, false); SET2(1, 0, false); SET2(0, 1, false); SET2(1, 1, false); SET2(0, 2, false); SET2(1, 2, false);
  SET2(0, 1, true); SET2(1, 1, true); SET2(0, 2, true); SET2(1, 2, true);
#undef SET2
  return e;
}
static int set_kernel_attrs(edgpu_ctx *c) {
  if (c->fast_attrs_set) return EDGPU_OK;
  CK(cudaSetDevice(c->device));
  CK((set_fcol_attr<8, 0>())); CK((set_fcol_attr<8, 1>())); CK((set_fcol_attr<8, 2>()));
  CK((set_fcol_attr<12, 0>())); CK((set_fcol_attr<12, 1>())); CK((set_fcol_attr<12, 2>()));
  CK((set_fcol_attr<16, 0>())); CK((set_fcol_attr<16, 1>())); CK((set_fcol_attr<16, 2>()));
  CK(set_fcol2_attr<8>()); CK(set_fcol2_attr<12>()); CK(set_fcol2_attr<16>());
  CK(cudaFuncSetAttribute(k_halo_push, cudaFuncAttributeMaxDynamicSharedMemorySize, PUSH_SMEM));
  CK(cudaFuncSetAttribute(k_srow<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
  CK(cudaFuncSetAttribute(k_srow<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
  CK(cudaFuncSetAttribute(k_srow<5, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
  CK(cudaFuncSetAttribute(k_srow<5, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
  c->fast_attrs_set = true;
  return EDGPU_OK;
}

int fast_plan_build(edgpu_ctx *c) {
  if (c->fplan) return EDGPU_OK;
  TRY(set_kernel_attrs(c));
  FastPlan *p = new FastPlan();
  c->fplan = p;
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
  int rc = build_srow(c, p->sr);
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

static size_t halo_buf_bytes(const edgpu_ctx *c, const SRowPlan &sr) {
  return (((size_t)sr.maxslot * (size_t)c->dimup + 2) * sizeof(double) + 255) & ~(size_t)255;
}
static double *halo_buf(const edgpu_ctx *c, const SRowPlan &sr, int p, unsigned long long epoch) {
  return reinterpret_cast<double *>(c->sym_peer[p] + EDGPU_HALO_HDR + (size_t)(epoch & 1ull) * halo_buf_bytes(c, sr));
}

template <int WT, int DIAG, int MODE, bool LISTS>
static void launch_fcol_m(int uni, int grid, size_t smem, cudaStream_t st, const FColArgs &a) {
  if (uni == 3) k_fcol<WT, DIAG, 3, MODE, LISTS><<<grid, FCOL_THREADS, smem, st>>>(a);
  else if (uni == 2) k_fcol<WT, DIAG, 2, MODE, LISTS><<<grid, FCOL_THREADS, smem, st>>>(a);
  else if (uni == 1) k_fcol<WT, DIAG, 1, MODE, LISTS><<<grid, FCOL_THREADS, smem, st>>>(a);
  else k_fcol<WT, DIAG, 0, MODE, LISTS><<<grid, FCOL_THREADS, smem, st>>>(a);
}
template <int WT, int DIAG>
static void launch_fcol(int uni, int mode, bool lists, int grid, size_t smem, cudaStream_t st, const FColArgs &a) {
  if constexpr (DIAG == 0) {
    if (lists) {
      if (mode == 2) launch_fcol_m<WT, 0, 2, true>(uni, grid, smem, st, a);
      else launch_fcol_m<WT, 0, 1, true>(uni, grid, smem, st, a);
      return;
    }
    if (mode == 2) { launch_fcol_m<WT, 0, 2, false>(uni, grid, smem, st, a); return; }
  }
  if (mode == 1) launch_fcol_m<WT, DIAG, 1, false>(uni, grid, smem, st, a);
  else launch_fcol_m<WT, DIAG, 0, false>(uni, grid, smem, st, a);
}
template <int DIAG>
static void launch_fcol_w(int WT, int uni, int mode, bool lists, int grid, size_t smem, cudaStream_t st, const FColArgs &a) {
  if (WT == 8) launch_fcol<8, DIAG>(uni, mode, lists, grid, smem, st, a);
  else if (WT == 12) launch_fcol<12, DIAG>(uni, mode, lists, grid, smem, st, a);
  else launch_fcol<16, DIAG>(uni, mode, lists, grid, smem, st, a);
}

template <int WT, int UNI>
static void launch_fcol2_u(int mode, bool lists, int grid, size_t smem, cudaStream_t st, const FColArgs &a) {
  if (lists) {
    if (mode == 2) k_fcol2<WT, UNI, 2, true><<<grid, FCOL_THREADS, smem, st>>>(a);
    else k_fcol2<WT, UNI, 1, true><<<grid, FCOL_THREADS, smem, st>>>(a);
  } else {
    if (mode == 2) k_fcol2<WT, UNI, 2, false><<<grid, FCOL_THREADS, smem, st>>>(a);
    else if (mode == 1) k_fcol2<WT, UNI, 1, false><<<grid, FCOL_THREADS, smem, st>>>(a);
    else k_fcol2<WT, UNI, 0, false><<<grid, FCOL_THREADS, smem, st>>>(a);
  }
}
static void launch_fcol2(int WT, int uni, int mode, bool lists, int grid, size_t smem, cudaStream_t st, const FColArgs &a) {
  if (WT == 8) { if (uni) launch_fcol2_u<8, 1>(mode, lists, grid, smem, st, a); else launch_fcol2_u<8, 0>(mode, lists, grid, smem, st, a); }
  else if (WT == 12) { if (uni) launch_fcol2_u<12, 1>(mode, lists, grid, smem, st, a); else launch_fcol2_u<12, 0>(mode, lists, grid, smem, st, a); }
  else { if (uni) launch_fcol2_u<16, 1>(mode, lists, grid, smem, st, a); else launch_fcol2_u<16, 0>(mode, lists, grid, smem, st, a); }
}

// y (+)= [Hd o x +] F_k x on a matrix whose contiguous dimension is factor k's index.  dw_lists: also apply the dw
// hops the row pass left to this one (sharded vector, k == 0, accumulate form).
int fast_apply_col(edgpu_ctx *c, int k, bool with_diag, bool acc, const double *d_x, double *d_y, int64_t ncols, int64_t coloff,
                   double *d_xp, int *npartials, bool dw_lists, int64_t list_col0, int grid_limit) {
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
  int grid = (int)std::min<int64_t>(ncols, c->sm_count);
  if (grid_limit > 0 && grid > grid_limit) grid = grid_limit;
  if (grid < 1) return EDGPU_OK;
  const int diag = !with_diag ? 0 : (c->d_diag ? 1 : 2);
  const int mode = d_xp ? 2 : (acc ? 1 : 0);
  if (mode == 2 && (diag != 0 || !acc)) return edgpu_set_err(EDGPU_ERR_INVALID, "Lanczos epilogue needs the accumulate form without diagonal");
  const bool lists = dw_lists && p->sr.lists;
  if (lists) {
    if (k != 0 || diag != 0 || !acc) return edgpu_set_err(EDGPU_ERR_INVALID, "dw source lists belong to the accumulating up pass");
    const SRowPlan &sr = p->sr;
    a.lptr = sr.d_lptr + list_col0; a.lown = sr.d_lown; a.lcol = sr.d_lcol2; a.lamp = sr.d_lamp;
    const double *hb = halo_buf(c, sr, c->rank, c->halo_epoch);
    for (int q = 0; q < EDGPU_MAXP; q++) a.xb[q] = hb;
    a.xb[c->rank] = d_x - list_col0 * c->dimup;                      // own cut groups: columns of the local shard
    a.lflag = sr.d_lflag + list_col0;
    a.diagmode = c->d_diag ? 1 : 2;
    if (c->d_diag) a.diag = c->d_diag + list_col0 * c->dimup;      // the skipped columns' stored diagonal, same column window
  }
  // *npartials on entry = partial sums already in c->d_partials (earlier column windows of the same H*v)
  const int pbase = (npartials && d_xp) ? *npartials : 0;
  a.xp = d_xp; a.st = c->d_st; a.partials = c->d_partials + pbase;
  a.vect = (d_xp && c->sweep_vect) ? c->sweep_vect + list_col0 * c->dimup : nullptr;
  if (npartials) *npartials = pbase + grid;
  if (use2) {
    if (diag != 0) return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "cluster column kernel has no fused diagonal");
    int pairs = (int)std::min<int64_t>(ncols, c->sm_count / 2);
    if (grid_limit > 0 && 2 * pairs > grid_limit) pairs = std::max(1, grid_limit / 2);
    const int uni2 = (ff.uniform && c->opt_no_uniform != 1) ? 1 : 0;
    if (npartials) *npartials = pbase + 2 * pairs;
    launch_fcol2(ff.WT, uni2, mode, lists, 2 * pairs, p->col2_smem[k], c->stream, a);
    CKL(c);
    return EDGPU_OK;
  }
  // no_uniform: 0 = best available, 1 = force the value-table kernel, 2 = 4-byte entries (uniform kernel when it applies)
  int uni = 0;
  if (ff.uniform && c->opt_no_uniform != 1) uni = (ff.d_ell16 && c->opt_no_uniform != 2) ? 2 : 1;
  else if (ff.korder && ff.d_ell16 && c->opt_no_uniform == 0) uni = 3;
  a.ns = c->ns;
  for (int q = 0; q < EDGPU_MAX_SITES; q++) a.vk[q] = ff.vk[q];
  if (diag == 0) launch_fcol_w<0>(ff.WT, uni, mode, lists, grid, p->col_smem[k], c->stream, a);
  else if (diag == 1) launch_fcol_w<1>(ff.WT, uni, mode, false, grid, p->col_smem[k], c->stream, a);
  else launch_fcol_w<2>(ff.WT, uni, mode, false, grid, p->col_smem[k], c->stream, a);
  CKL(c);
  return EDGPU_OK;
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

// y = Hd o x + x Hdw^T on the whole low groups of the local shard (every source local).  grid_limit > 0: leave the
// other SMs to a kernel that runs next to this one (the halo copy).
int fast_apply_row(edgpu_ctx *c, const double *d_x, double *d_y, int grid_limit) {
  TRY(fast_plan_build(c));
  const SRowPlan &sr = c->fplan->sr;
  if (!sr.ok) return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "structured row kernel does not cover this model");
  if ((reinterpret_cast<uintptr_t>(d_x) & 15) != 0) return edgpu_set_err(EDGPU_ERR_INVALID, "fast H*v needs 16-byte aligned vectors");
  if (sr.nchunks == 0 || c->qdw == 0) return EDGPU_OK;            // this rank owns no whole group: all left to the lists
  SRowArgs a{};
  a.x = d_x; a.y = d_y; a.n = (int)c->dimup; a.n8 = (uint32_t)(c->dimup * 8);
  a.nchunks = sr.nchunks; a.tbits = sr.T; a.nhigh = sr.nhigh; a.cpad = sr.cmax;
  a.chunks = sr.d_chunks; a.recs = sr.d_recs;
  for (int k = 0; k < EDGPU_MAX_SITES; k++) a.vk[k] = sr.vk[k];
  a.vkd = sr.d_vk;
  a.diag = c->d_diag; a.dr0 = sr.d_dr0; a.dr1 = sr.d_dr1; a.dfac_s = c->dw.d_dfac;
  a.c0 = (int)c->coloff;
  const int diag = c->d_diag ? 1 : 2;
  const int64_t nitems = ((c->dimup + SROW_R - 1) / SROW_R) * sr.nchunks;
  int grid = (int)std::min<int64_t>(nitems, c->sm_count);
  if (grid_limit > 0 && grid > grid_limit) grid = grid_limit;
  CUtensorMap tmx;
  TRY(make_tmap_2d(&tmx, d_x, (uint64_t)c->dimup, (uint64_t)c->qdw, SROW_R, SROW_BC));
  if (sr.LR == 4) {
    if (diag == 1) k_srow<4, 1><<<grid, SROW_THREADS, sr.smem, c->stream>>>(tmx, a);
    else k_srow<4, 2><<<grid, SROW_THREADS, sr.smem, c->stream>>>(tmx, a);
  } else {
    if (diag == 1) k_srow<5, 1><<<grid, SROW_THREADS, sr.smem, c->stream>>>(tmx, a);
    else k_srow<5, 2><<<grid, SROW_THREADS, sr.smem, c->stream>>>(tmx, a);
  }
  CKL(c);
  return EDGPU_OK;
}

// ---- halo plumbing (sharded vector) ------------------------------------------------------------------------------
// slab layout, identical on every rank: [EDGPU_HALO_HDR bytes of arrival flags][buffer 0][buffer 1], a buffer =
// maxslot columns of DimUp doubles
int fast_halo_bytes(edgpu_ctx *c, size_t *bytes) {
  *bytes = 0;
  if (c->nranks <= 1 || c->nranks > EDGPU_MAXP) return EDGPU_OK;
  if (!fast_supported_local(c)) return EDGPU_OK;                 // builds the plan (and the tables) on first use
  const SRowPlan &sr = c->fplan->sr;
  if (!sr.lists) return EDGPU_OK;
  *bytes = EDGPU_HALO_HDR + 2 * halo_buf_bytes(c, sr);
  return EDGPU_OK;
}
bool fast_peer_ready(edgpu_ctx *c) {
  return c->nranks > 1 && c->nranks <= EDGPU_MAXP && c->sym_ok && c->sym_slab != nullptr;
}
// this rank's source columns -> the halo buffers (epoch c->halo_epoch) of the ranks that list them; w = column window
// of the targets (its own launch, counter and arrival flags), w < 0: every window one after another
int fast_halo_push(edgpu_ctx *c, const double *d_x, cudaStream_t st, int ctas, int w) {
  const SRowPlan &sr = c->fplan->sr;
  if (w < 0) {
    for (int k = 0; k < sr.nwin; k++) TRY(fast_halo_push(c, d_x, st, ctas, k));
    return EDGPU_OK;
  }
  PushArgs a{};
  a.x = d_x; a.n = (int)c->dimup; a.npush = sr.pwin[w + 1] - sr.pwin[w];
  a.pdst = sr.d_pdst + sr.pwin[w]; a.pslot = sr.d_pslot + sr.pwin[w]; a.psrc = sr.d_psrc + sr.pwin[w];
  for (int p = 0; p < EDGPU_MAXP; p++) {
    const int q = p < c->nranks ? p : c->rank;
    a.hb[p] = halo_buf(c, sr, q, c->halo_epoch);
    a.flag[p] = (p < c->nranks && p != c->rank && !c->halo_emul)
                    ? reinterpret_cast<unsigned long long *>(c->sym_peer[p]) + w * EDGPU_MAXP + c->rank : nullptr;
  }
  a.epoch = c->halo_epoch; a.ctr = sr.d_pushctr + w;
  const int64_t nitems = (int64_t)a.npush * ((c->dimup * 8 + PUSH_SB - 1) / PUSH_SB);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(std::max<int64_t>(nitems, 1), ctas));   // runs even without work: flags
  k_halo_push<<<grid, 32, PUSH_SMEM, st>>>(a);
  CKL(c);
  return EDGPU_OK;
}

// d_xp != nullptr: Lanczos form -- d_y only holds the row-pass partial result, w = sx*(H x) - cprev*xp goes to
// d_xp and the per-CTA partial sums of (sx*x).w to c->d_partials (*npartials of them)
//
// Sharded pipeline (no cross-rank barrier, no remote reads):
//   stream2 : k_halo_push window 1 | window 2 | ... | window K   (this rank's columns -> the peers' halo buffers of
//             this epoch, ordered by the column window of the TARGET, each followed by its arrival flags)
//   stream  : k_srow (local groups, the other SMs) | k_halo_wait 1 | k_fcol window 1 (+ arrived columns) | ... | K
// Ordering: a peer overwrites buffer b = epoch & 1 again at epoch + 2, which it starts only after it has seen this
// rank's flag of epoch + 1, and this rank raises that flag in stream order after its k_fcol of this epoch.
int fast_apply_local(edgpu_ctx *c, const double *d_x, double *d_y, double *d_xp, int *npartials) {
  TRY(fast_plan_build(c));
  const SRowPlan &sr = c->fplan->sr;
  if (npartials) *npartials = 0;
  if (c->nranks == 1) {
    // row tiles first (y = Hd o x + x Hdw^T, write only), then whole columns (y += Hup x, contiguous RMW)
    prof_mark(c, "k_srow");
    TRY(fast_apply_row(c, d_x, d_y, 0));
    prof_mark(c, "k_fcol");
    return fast_apply_col(c, 0, false, true, d_x, d_y, c->qdw, c->coloff, d_xp, npartials, true, 0, 0);
  }
  if (!fast_peer_ready(c)) return edgpu_set_err(EDGPU_ERR_INVALID, "sharded fast H*v: no halo slab (peers not mapped)");
  const unsigned long long *flags = reinterpret_cast<const unsigned long long *>(c->sym_slab);
  if (c->halo_emul) {                                              // emulation: the selftest has pushed for every rank
    prof_mark(c, "k_srow");
    TRY(fast_apply_row(c, d_x, d_y, 0));
    prof_mark(c, "k_fcol");
    return fast_apply_col(c, 0, false, true, d_x, d_y, c->qdw, c->coloff, d_xp, npartials, true, 0, 0);
  }
  c->halo_epoch++;
  const int K = sr.nwin;
  if (!c->stream2 || c->opt_no_overlap) {
    prof_mark(c, "k_halo_push");
    TRY(fast_halo_push(c, d_x, c->stream, c->sm_count, -1));
    prof_mark(c, "k_srow");
    TRY(fast_apply_row(c, d_x, d_y, 0));
    prof_mark(c, "k_halo_wait");
    for (int w = 0; w < K; w++) {
      k_halo_wait<<<1, 32, 0, c->stream>>>(flags + w * EDGPU_MAXP, c->nranks, c->rank, c->halo_epoch);
      CKL(c);
    }
    prof_mark(c, "k_fcol");
    return fast_apply_col(c, 0, false, true, d_x, d_y, c->qdw, c->coloff, d_xp, npartials, true, 0, 0);
  }
  // CTAs (= SMs) of the push next to the row pass, measured on C3: at 8 ranks the push is the critical path and the row
  // pass short (64 CTAs 0.88 ms per H*v, 32 CTAs 0.96 ms); at 4 ranks the row pass still needs its SMs (64 CTAs 1.30 ms,
  // 32 CTAs 1.14 ms); at 2 the two are balanced at 32
  const int hctas = (int)std::max<int64_t>(1, std::min<int64_t>(c->opt_halo_ctas > 0 ? c->opt_halo_ctas : (c->nranks >= 8 ? 64 : 32), c->sm_count / 2));
  CK(cudaEventRecord(c->ev_fork, c->stream));
  CK(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
  for (int w = 0; w < K; w++) TRY(fast_halo_push(c, d_x, c->stream2, hctas, w));
  CK(cudaEventRecord(c->ev_join, c->stream2));
  prof_mark(c, "k_srow");
  TRY(fast_apply_row(c, d_x, d_y, c->sm_count - hctas));
  prof_mark(c, "k_fcol");                                          // includes the waits for the arrivals
  for (int w = 0; w < K; w++) {
    const int64_t j0 = c->qdw * w / K, j1 = c->qdw * (w + 1) / K;
    k_halo_wait<<<1, 32, 0, c->stream>>>(flags + w * EDGPU_MAXP, c->nranks, c->rank, c->halo_epoch);
    CKL(c);
    // the caller may overwrite d_x as soon as this stream is done: its own pushes (which read d_x) end before that
    if (w == K - 1) CK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    if (j1 <= j0) continue;
    TRY(fast_apply_col(c, 0, false, true, d_x + j0 * c->dimup, d_y + j0 * c->dimup, j1 - j0, c->coloff + j0,
                       d_xp ? d_xp + j0 * c->dimup : nullptr, npartials, true, j0, w + 1 < K ? c->sm_count - hctas : 0));
  }
  return EDGPU_OK;
}

extern "C" int edgpu_halo_info(edgpu_ctx *c, int64_t *bytes_out, int64_t *bytes_in, int *windows) {
  if (!c || !c->hstatus) return edgpu_set_err(EDGPU_ERR_INVALID, "halo_info: no live sector");
  if (bytes_out) *bytes_out = 0;
  if (bytes_in) *bytes_in = 0;
  if (windows) *windows = 0;
  if (c->nranks == 1 || !fast_peer_ready(c) || !c->fplan || !c->fplan->sr.lists) return EDGPU_OK;
  const SRowPlan &sr = c->fplan->sr;
  if (bytes_out) *bytes_out = (int64_t)sr.npush * c->dimup * 8;
  if (bytes_in) *bytes_in = (int64_t)sr.nslot * c->dimup * 8;
  if (windows) *windows = sr.nwin;
  return EDGPU_OK;
}
