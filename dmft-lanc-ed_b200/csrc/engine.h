// engine.h -- internal state of the B200 engine (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/edgpu.h"
#include "hd_funcs.h"

struct Factor {                 // spH0ups(1) / spH0dws(1) on device, CSR in insertion order
  int64_t n = 0, nnz = 0;
  int maxrow = 0;
  int32_t *d_map = nullptr;     // Hs(.)%map
  int32_t *d_rowptr = nullptr;  // n+1
  int32_t *d_cols = nullptr;
  double *d_vals = nullptr;
  double *d_dfac = nullptr;     // factorised diagonal table (direct mode)
};

struct LancState {              // device-resident Lanczos scalars (no host sync inside a step)
  double red;                   // last reduction result (all-reduced in place with NCCL)
  double alpha, beta;
  double sx;                    // v_k     = sx * X
  double cprev;                 // beta_{k-1} * (scale of Xp)
  double norm2;
  double sw_a, sw_zk;           // second sweep of sp_lanc_eigh (k_fcol's epilogue when a vect pointer is passed): a_k, Z(k,1)
};

#define EDGPU_MAXP 8            // ranks of one NVLink domain (peer-mapped symmetric slab)
#define EDGPU_MAX_WINDOWS 8     // column windows of the sharded H*v pipeline (push of window w+1 under the column pass of w)
#define EDGPU_HALO_HDR 1024     // bytes of arrival flags in front of the halo buffers: [window][source rank] 64-bit epochs

// One low group of the structured row kernel (hxv_fast.cu), precomputed per sector: 80 bytes, bulk-copied next
// to the tile.  hx = high word h | (mask of the high-bit hops the kernel applies) << 16; par bit kk = parity of
// the occupied high bits below kk; pc[kk] = first column of the partner group of hop kk (tile-local for the T
// lowest high bits, shard-local otherwise).
struct SRowRec {
  int32_t lb, N;
  uint32_t hx, par;
  int32_t pc[16];
};
static_assert(sizeof(SRowRec) == 80, "SRowRec is copied in 16-byte units");

struct TiledPlan;               // hxv_tiled.cu
struct FastPlan;                // hxv_fast.cu
struct OrbsPlan;                // orbs.cu

struct edgpu_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;             // halo copy next to the row kernel (sharded fast path)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_fork = nullptr, ev_join = nullptr;   // fork / join of stream2 around an H*v
  bool fast_attrs_set = false, tiled_attrs_set = false;   // per-device function attributes (this context's device)
  int sm_count = 148;
  // inputs
  edgpu_params hp{};
  std::vector<double> h_hloc, h_be, h_bv, h_bh;
  DevParams dp{};
  int ns = 0;
  uint32_t h_binom[EDGPU_BINOM_LD * EDGPU_BINOM_LD];
  uint32_t *d_binom = nullptr;
  // communicator
  int rank = 0, nranks = 1;
  void *comm = nullptr;         // ncclComm_t
  // live sector (Hstatus / Hsector, ED_HAMILTONIAN_COMMON.f90:17-18)
  bool hstatus = false;
  int isector = 0, nup = 0, ndw = 0;
  int64_t dimup = 0, dimdw = 0;
  int64_t qdw = 0, coloff = 0, nloc = 0;      // dw split (columns owned); nloc = vecDim_Hv_sector = nel * dimph
  int64_t nel = 0;                            // electron part of the local vector: DimUp * mpiQdw
  int dimph = 1;                              // DimPh = Nph + 1 phonon slabs (phonon index slowest, ED_SETUP.f90:133)
  int64_t qup = 0, rowoff = 0;                // up split used by the transposed layout
  Factor up, dw;
  double *d_diag = nullptr;                   // spH0d (stored mode), nloc values
  int64_t *d_nd_rowptr = nullptr, *d_nd_cols = nullptr;
  double *d_nd_vals = nullptr;
  int64_t nd_nnz = 0;
  // work buffers (grown on demand, freed by delete_hv_sector)
  double *d_in = nullptr, *d_out = nullptr;   // staging for the host-pointer operator
  double *d_vt = nullptr, *d_hvt = nullptr;   // transposed shard DimDw x qup
  double *d_send = nullptr, *d_recv = nullptr;
  double *d_full = nullptr;                   // all-gathered vector for spH0nd when nranks>1
  double *d_lx = nullptr, *d_lp = nullptr, *d_lt = nullptr, *d_l0 = nullptr, *d_lv = nullptr;
  double *d_partials = nullptr;
  LancState *d_st = nullptr;
  double *d_alanc = nullptr, *d_blanc = nullptr;
  int lanc_cap = 0;
  double *h_pinned = nullptr;                 // small pinned scratch for scalar read-back
  // GF state
  double *d_gs = nullptr;
  int gs_nup = -1, gs_ndw = -1;
  int64_t gs_nloc = 0;
  double gs_e0 = 0.0;
  double *sweep_vect = nullptr;               // non-null only while the second sweep of sp_lanc_eigh runs: vect += zk * v_k in k_fcol's epilogue
  bool lv_valid = false;                      // d_lv holds the normalised eigenvector of the last sp_lanc_eigh (live sector)
  // options
  int algo = EDGPU_ALGO_AUTO;
  int64_t opt_tile_rows = 0, opt_tile_h = -1, opt_col_h = -1;
  TiledPlan *plan = nullptr;
  FastPlan *fplan = nullptr;
  OrbsPlan *orbs = nullptr;                   // live sector of an ed_total_ud = F model (orbs.cu); then up / dw are empty
  int64_t opt_srow_lr = 0, opt_srow_t = 0, opt_no_uniform = 0, opt_no_fuse = 0, opt_no_peer = 0, opt_col_cluster = 0;
  int64_t opt_halo_ctas = 0, opt_no_overlap = 0, opt_no_batch = 0, opt_halo_windows = 0;
  int64_t launches = 0;
  // symmetric slab (nranks > 1): one allocation per rank with the same layout, opened by every peer through CUDA
  // IPC.  It holds the halo of the sharded fast path: a block of arrival flags and two buffers of source columns
  // that the OWNERS store into over NVLink (k_halo_push); nobody ever reads a peer's memory.
  char *sym_slab = nullptr;
  size_t sym_bytes = 0;
  char *sym_peer[64] = {nullptr};
  bool sym_ok = false;
  bool halo_emul = false;                     // selftest only: ranks emulated on one device (sym_peer set by hand, no flags)
  unsigned long long halo_epoch = 0;          // H*v counter of the sharded fast path: buffer = epoch & 1, flag value = epoch
  // per-pass timing (edgpu_time_hxv_passes): events recorded between the kernels of one H*v
  bool prof = false;
  int prof_n = 0;
  cudaEvent_t pev[8] = {nullptr};
  const char *prof_name[7] = {nullptr};
};

// marks the start of pass `name` (and the end of the previous one) when per-pass timing is on
static inline void prof_mark(edgpu_ctx *c, const char *name) {
  if (!c->prof || c->prof_n >= 7) return;
  cudaEventRecord(c->pev[c->prof_n], c->stream);
  c->prof_name[c->prof_n] = name;
  c->prof_n++;
}

// ---- error plumbing -------------------------------------------------------------------------
int edgpu_set_err(int code, const char *fmt, ...);
#define CK(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return edgpu_set_err(EDGPU_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,         \
                           cudaGetErrorString(e_));                                           \
  } while (0)
#define CKL(c)                                                                                \
  do {                                                                                        \
    (c)->launches++;                                                                          \
    cudaError_t e_ = cudaGetLastError();                                                      \
    if (e_ != cudaSuccess)                                                                    \
      return edgpu_set_err(EDGPU_ERR_CUDA, "%s:%d kernel launch: %s", __FILE__, __LINE__,     \
                           cudaGetErrorString(e_));                                           \
  } while (0)
#define TRY(expr)                                                                             \
  do {                                                                                        \
    int rc_ = (expr);                                                                         \
    if (rc_) return rc_;                                                                      \
  } while (0)

// ---- cross-file entry points ------------------------------------------------------------------
// hxv.cu: y = H x on the local shard (device pointers), all terms, any nranks
int hxv_apply(edgpu_ctx *c, const double *d_x, double *d_y);
// true when hxv_apply would take the fast two-kernel path for this vector (then the Lanczos epilogue can be fused)
bool hxv_fast_path(edgpu_ctx *c, const double *d_x);
// orbs.cu: ed_total_ud = F (one (Nup, Ndw) pair per orbital)
int orbs_build(edgpu_ctx *c, int isector);
int orbs_free(edgpu_ctx *c);
int orbs_apply(edgpu_ctx *c, const double *d_x, double *d_y);
void orbs_factor_host(const DevParams &dp, int f, int n, std::vector<int32_t> &map, std::vector<int32_t> &rowptr,
                      std::vector<int32_t> &cols, std::vector<double> &vals);   // host arithmetic of one factor
// hxv_tiled.cu
int tiled_plan_build(edgpu_ctx *c);
int tiled_plan_free(edgpu_ctx *c);
bool tiled_supported(const edgpu_ctx *c);
int tiled_apply_local(edgpu_ctx *c, const double *d_x, double *d_y);   // nranks==1: full operator
// y = [Hd o x +] F_k x, contiguous dimension = index of factor k (0 up, 1 dw), ncols local columns
int tiled_apply_col(edgpu_ctx *c, int k, bool with_diag, const double *d_x, double *d_y, int64_t ncols, int64_t coloff);
// hxv_fast.cu: host arithmetic of the structured row kernel's plan (no device needed; see selftest.cu)
struct SRowHostPlan {
  bool ok = false;
  int LR = 0, T = 0, nhigh = 0, cmax = 0, cpad = 0, maxg = 0;
  size_t smem = 0;
  std::vector<int> coloffs;            // [P+1] first global column of every rank
  std::vector<int32_t> jhi;            // [2^nhigh] first global column of group h, -1 = empty group
  std::vector<char> gwhole;            // [2^nhigh] every column of the group is on this rank
  std::vector<int4> chunks;            // (first record, groups, local column begin, local column end)
  std::vector<SRowRec> recs;
  std::vector<int> lptr, lown, lcol;   // source lists of the hops the row kernel leaves out (srow_lists_host)
  std::vector<double> lamp;
  std::vector<unsigned char> lflag;    // per local column: 1 = not written by the row kernel, 2 = has entries
  std::vector<int> zcols;              // local columns that have entries
};
// returns 1 = plan built, 0 = the structured kernel does not apply, -1 = internal inconsistency
int srow_plan_host(int ns, int ndw, int64_t dimdw, int nranks, int rank, int lr, int tbits_opt, SRowHostPlan &hp);
void srow_lists_host(SRowHostPlan &hp, int rank, int64_t dimdw, const int32_t *map, const int32_t *rp, const int32_t *cc, const double *vv);
// halo tables of the push model (pure host arithmetic): lcol2 = consumer column of every list entry of rank `me` (halo
// slot for a remote owner, local column for an own one); triples (pdst, pslot, psrc) = what `me` stores into its
// peers' halo buffers, sorted by (column window of the target, source column), pwin[K+1] = triple range per window
int halo_tables_host(int ns, int ndw, int64_t dimdw, int P, int me, int lr, int tbits_opt, const SRowHostPlan &hp_me,
                     const int32_t *map, const int32_t *rp, const int32_t *cc, const double *vv, int K, std::vector<int> &lcol2,
                     std::vector<int> &pdst, std::vector<int> &pslot, std::vector<int> &psrc, int *pwin, int *nslot, int *maxslot);
// hxv_fast.cu: TMA-staged whole-column kernel + structured single-band row kernel
int fast_plan_build(edgpu_ctx *c);
int fast_plan_free(edgpu_ctx *c);
bool fast_supported_local(edgpu_ctx *c);          // full operator on the local shard (pushed halo when nranks > 1)
bool fast_supported_col(edgpu_ctx *c, int k);     // whole-column kernel for factor k
int fast_apply_local(edgpu_ctx *c, const double *d_x, double *d_y, double *d_xp = nullptr, int *npartials = nullptr);
int fast_apply_col(edgpu_ctx *c, int k, bool with_diag, bool acc, const double *d_x, double *d_y, int64_t ncols, int64_t coloff,
                   double *d_xp = nullptr, int *npartials = nullptr, bool dw_lists = false, int64_t list_col0 = 0, int grid_limit = 0);
int fast_apply_row(edgpu_ctx *c, const double *d_x, double *d_y, int grid_limit = 0);
bool fast_peer_ready(edgpu_ctx *c);                // sharded: the halo slab exists and every peer's copy is mapped
int fast_halo_bytes(edgpu_ctx *c, size_t *bytes);   // size of the halo slab of the live sector (0: no sharded fast path)
// this rank's columns -> the peers' halo buffers, column window w of the targets (w < 0: every window)
int fast_halo_push(edgpu_ctx *c, const double *d_x, cudaStream_t st, int ctas, int w);
// comm.cu
int comm_symm_setup(edgpu_ctx *c, size_t bytes);        // collective: `bytes` per rank, mapped by every peer
int comm_symm_teardown(edgpu_ctx *c);                   // collective
int comm_barrier(edgpu_ctx *c);                         // stream-ordered cross-rank barrier (tiny all-reduce)
int vec_alloc(edgpu_ctx *c, double **p, int64_t n);     // engine work vector of n doubles (+ padding for 16-byte copies)
void vec_free(edgpu_ctx *c, double **p);
int comm_allreduce_scalar(edgpu_ctx *c, double *d_scalar);
int comm_allreduce_array(edgpu_ctx *c, double *d_a, int n);
// capi.cu
int build_sector_map_device(edgpu_ctx *c, int npart, int32_t **d_map);
int comm_transpose_fwd(edgpu_ctx *c, const double *d_x, double *d_vt);          // V(DimUp,qdw) -> Vt(DimDw,qup)
int comm_transpose_bwd_add(edgpu_ctx *c, const double *d_hvt, double *d_y);     // Hv += (Hvt)^T
int comm_allgather(edgpu_ctx *c, const double *d_x, double *d_full);
// grouped point-to-point exchange of doubles: segment p of send (soff[p], scnt[p]) goes to rank p, segment p of recv
// comes from rank p; the own segment is a device copy
int comm_exchange(edgpu_ctx *c, const double *send, const int64_t *soff, const int64_t *scnt, double *recv,
                  const int64_t *roff, const int64_t *rcnt);
