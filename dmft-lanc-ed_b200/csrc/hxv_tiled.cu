// hxv_tiled.cu -- shared-memory staged H*v for large sectors (the engine's fast path).
//
// The sector vector is the column-major matrix x(i_up, i_dw).  y = Hd o x + Hup x + x Hdw^T is done
// in two passes, each of which keeps a tile of x in shared memory and applies ONE one-spin factor to
// it, so that the Ns/2 gathers per element and factor hit shared memory instead of L2/HBM:
//
//   pass 1  k_tile_col : tile = (contiguous chunk of i_up rows) x (G columns); y = Hd o x + Hup x
//   pass 2  k_tile_row : tile = (R consecutive i_up rows) x (contiguous chunk of i_dw columns);
//                        y += x Hdw^T  (lanes run along the contiguous i_up index, so both the global
//                        traffic and the shared-memory gathers are unit-stride)
//
// The factors spH0ups(1)/spH0dws(1) (ED_HAMILTONIAN/stored/H_up.f90, H_dw.f90) are repacked once per
// sector into a 4-byte ELL entry (column | value-id | sign); each CTA keeps the slice of its chunk in
// shared memory for its whole life.  A hop whose source lies outside the chunk is read from global
// memory (L2).  Chunks are plain contiguous index ranges of the sorted basis, so nothing here depends
// on the bath geometry; the chunk sizes are chosen from the shared-memory budget of 2 CTAs per SM.
//
// HBM traffic per element: pass 1 reads x, writes y (16 B, +8 B when spH0d is streamed);
// pass 2 reads x and y, writes y (24 B).  See DESIGN.md for the roofline discussion.
#include <algorithm>
#include <map>
#include <vector>

#include "engine.h"

#define ELL_COL_BITS 20
#define ELL_COL_MASK 0xFFFFFu
#define ELL_VID_MASK 0x7FFu
#define TILED_MAX_VALS 256
#define TILED_MAX_W 16
#define TILED_THREADS 512

struct PackedFactor {
  int W = 0, nvals = 0;
  int64_t n = 0;
  uint32_t *d_ell = nullptr;    // [W][n], slot-major
  double *d_vtab = nullptr;     // [nvals], vtab[0] = 0
};

struct TiledPlan {
  PackedFactor pf[2];           // 0 = up, 1 = dw
  // pass 1 (column kernel on the up factor) / transposed dw apply for the sharded path
  int a_nchunks[2] = {0, 0}, a_G[2] = {1, 1};
  int64_t a_chunk[2] = {0, 0};
  size_t a_smem[2] = {0, 0};
  // pass 2 (row-tile kernel on the dw factor)
  int b_nchunks = 0, b_R = 16;
  int64_t b_chunk = 0;
  size_t b_smem = 0;
};

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

struct ColArgs {
  const double *x;
  double *y;
  int64_t n;                    // factor (contiguous) dimension
  int64_t ncols;                // columns of the local matrix
  int64_t coloff;               // global index of local column 0 (diagonal tables of the other spin)
  const uint32_t *ell;
  const double *vtab;
  int W, nvals;
  int64_t chunk;                // rows per chunk (even), nchunks = ceil(n/chunk)
  int nchunks, cpc;             // CTAs per chunk
  // diagonal
  const double *diag;           // DIAG==1
  const double *dfac_c, *dfac_s;// DIAG==2
  const int32_t *map_c, *map_s;
  int norb;
  double uloc[EDGPU_MAX_ORB];
  double ust;
};

// pass 1: y[:, j] = Hd o x[:, j] + F x[:, j] for G columns at a time, rows of one chunk per CTA
template <int G, int DIAG>
__global__ void __launch_bounds__(TILED_THREADS, 2) k_tile_col(ColArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int chunk_id = blockIdx.x % a.nchunks;
  const int lane_id = blockIdx.x / a.nchunks;
  const int64_t r0 = (int64_t)chunk_id * a.chunk;
  const int nr = (int)min(a.chunk, a.n - r0);
  const int ldt = (int)a.chunk;                                  // tile leading dimension
  double *tile = reinterpret_cast<double *>(smem_raw);           // [G][chunk]
  double *dtab = tile + (size_t)G * ldt;                         // [chunk]     (DIAG==2)
  double *vtab = dtab + (DIAG == 2 ? ldt : 0);                   // [TILED_MAX_VALS]
  double *colinfo = vtab + TILED_MAX_VALS;                       // [G][1+MAX_ORB]
  uint32_t *ell = reinterpret_cast<uint32_t *>(colinfo + G * (1 + EDGPU_MAX_ORB));   // [W][chunk]
  uint32_t *imp = ell + (size_t)a.W * ldt;                       // [chunk]     (DIAG==2)

  for (int i = threadIdx.x; i < a.nvals; i += blockDim.x) vtab[i] = a.vtab[i];
  for (int s = 0; s < a.W; s++)
    for (int r = threadIdx.x; r < nr; r += blockDim.x) ell[s * ldt + r] = a.ell[(size_t)s * a.n + r0 + r];
  if (DIAG == 2)
    for (int r = threadIdx.x; r < nr; r += blockDim.x) {
      dtab[r] = a.dfac_c[r0 + r];
      imp[r] = (uint32_t)a.map_c[r0 + r] & ((1u << a.norb) - 1u);
    }
  __syncthreads();

  const bool al16 = ((a.n & 1) == 0) && ((r0 & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.x) & 15) == 0);
  const int64_t ngroups = (a.ncols + G - 1) / G;
  for (int64_t cg = lane_id; cg < ngroups; cg += a.cpc) {
    const int64_t j0 = cg * G;
    // ---- stage the tile: G column segments of nr rows ----
#pragma unroll
    for (int g = 0; g < G; g++) {
      if (j0 + g >= a.ncols) break;
      const double *src = a.x + (j0 + g) * a.n + r0;
      double *dst = tile + (size_t)g * ldt;
      if (al16) {
        const int n2 = nr >> 1;
        for (int i = threadIdx.x; i < n2; i += blockDim.x) cp_async16(dst + 2 * i, src + 2 * i);
        if ((nr & 1) && threadIdx.x == 0) cp_async8(dst + nr - 1, src + nr - 1);
      } else {
        for (int i = threadIdx.x; i < nr; i += blockDim.x) cp_async8(dst + i, src + i);
      }
    }
    if (DIAG == 2 && threadIdx.x < G && j0 + threadIdx.x < a.ncols) {
      const int g = threadIdx.x;
      const int64_t jg = a.coloff + j0 + g;
      const uint32_t ms = (uint32_t)a.map_s[jg];
      colinfo[g * (1 + EDGPU_MAX_ORB)] = a.dfac_s[jg];
      for (int o = 0; o < a.norb; o++) {                         // X_o = sum_b W_ob n_dw,b
        double xo = 0.0;
        for (int b = 0; b < a.norb; b++)
          if ((ms >> b) & 1u) xo += (o == b) ? a.uloc[o] : a.ust;
        colinfo[g * (1 + EDGPU_MAX_ORB) + 1 + o] = xo;
      }
    }
    cp_async_wait_all();
    __syncthreads();
    // ---- compute ----
    for (int r = threadIdx.x; r < nr; r += blockDim.x) {
      double acc[G];
#pragma unroll
      for (int g = 0; g < G; g++) {
        double d = 0.0;
        if (DIAG == 1) d = (j0 + g < a.ncols) ? a.diag[(j0 + g) * a.n + r0 + r] : 0.0;
        if (DIAG == 2) {
          d = dtab[r] + colinfo[g * (1 + EDGPU_MAX_ORB)];
          const uint32_t mc = imp[r];
          for (int o = 0; o < a.norb; o++)
            if ((mc >> o) & 1u) d += colinfo[g * (1 + EDGPU_MAX_ORB) + 1 + o];
        }
        acc[g] = d * tile[(size_t)g * ldt + r];
      }
      for (int s = 0; s < a.W; s++) {
        const uint32_t e = ell[s * ldt + r];
        const uint32_t col = e & ELL_COL_MASK;
        double val = vtab[(e >> ELL_COL_BITS) & ELL_VID_MASK];
        if (e >> 31) val = -val;
        const uint32_t lc = col - (uint32_t)r0;
        if (lc < (uint32_t)nr) {
#pragma unroll
          for (int g = 0; g < G; g++) acc[g] += val * tile[(size_t)g * ldt + lc];
        } else {
#pragma unroll
          for (int g = 0; g < G; g++)
            if (j0 + g < a.ncols) acc[g] += val * __ldg(a.x + (j0 + g) * a.n + col);
        }
      }
#pragma unroll
      for (int g = 0; g < G; g++)
        if (j0 + g < a.ncols) a.y[(j0 + g) * a.n + r0 + r] = acc[g];
    }
    __syncthreads();
  }
}

struct RowArgs {
  const double *x;
  double *y;
  int64_t n;                    // contiguous dimension (i_up rows)
  int64_t nf;                   // factor dimension (i_dw columns, all local)
  const uint32_t *ell;
  const double *vtab;
  int W, nvals;
  int64_t chunk;                // columns per chunk
  int nchunks, cpc;
};

// pass 2: y[i, j] += sum_j' F(j, j') x[i, j'] for R consecutive rows i and one chunk of columns j
template <int R>
__global__ void __launch_bounds__(TILED_THREADS, 2) k_tile_row(RowArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int chunk_id = blockIdx.x % a.nchunks;
  const int lane_id = blockIdx.x / a.nchunks;
  const int64_t c0 = (int64_t)chunk_id * a.chunk;
  const int nc = (int)min(a.chunk, a.nf - c0);
  const int ldc = (int)a.chunk;
  double *tile = reinterpret_cast<double *>(smem_raw);           // [chunk][R]
  double *vtab = tile + (size_t)ldc * R;
  uint32_t *ell = reinterpret_cast<uint32_t *>(vtab + TILED_MAX_VALS);   // [W][chunk]

  for (int i = threadIdx.x; i < a.nvals; i += blockDim.x) vtab[i] = a.vtab[i];
  for (int s = 0; s < a.W; s++)
    for (int c = threadIdx.x; c < nc; c += blockDim.x) ell[s * ldc + c] = a.ell[(size_t)s * a.nf + c0 + c];
  __syncthreads();

  const int r = threadIdx.x % R;
  const int cl = threadIdx.x / R;
  const int cstep = blockDim.x / R;
  const int64_t ntiles = (a.n + R - 1) / R;
  for (int64_t t = lane_id; t < ntiles; t += a.cpc) {
    const int64_t i0 = t * R;
    const bool rok = (i0 + r) < a.n;
    for (int c = cl; c < nc; c += cstep)
      if (rok) cp_async8(tile + (size_t)c * R + r, a.x + (c0 + c) * a.n + i0 + r);
    cp_async_wait_all();
    __syncthreads();
    if (rok) {
      for (int c = cl; c < nc; c += cstep) {
        double *yp = a.y + (c0 + c) * a.n + i0 + r;
        double acc = *yp;
        for (int s = 0; s < a.W; s++) {
          const uint32_t e = ell[s * ldc + c];
          const uint32_t col = e & ELL_COL_MASK;
          double val = vtab[(e >> ELL_COL_BITS) & ELL_VID_MASK];
          if (e >> 31) val = -val;
          const uint32_t lc = col - (uint32_t)c0;
          if (lc < (uint32_t)nc) acc += val * tile[(size_t)lc * R + r];
          else acc += val * __ldg(a.x + (int64_t)col * a.n + i0 + r);
        }
        *yp = acc;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// plan: repack the factors, choose chunk sizes from the shared-memory budget
// ---------------------------------------------------------------------------------------------
static const size_t SMEM_BUDGET = 112 * 1024;          // 2 CTAs per SM (227 KB usable per SM)

static int pack_factor(edgpu_ctx *c, const Factor &f, PackedFactor &pf) {
  std::vector<int32_t> rp((size_t)f.n + 1), cols((size_t)std::max<int64_t>(f.nnz, 1));
  std::vector<double> vals((size_t)std::max<int64_t>(f.nnz, 1));
  CK(cudaMemcpy(rp.data(), f.d_rowptr, rp.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
  if (f.nnz) {
    CK(cudaMemcpy(cols.data(), f.d_cols, (size_t)f.nnz * sizeof(int32_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(vals.data(), f.d_vals, (size_t)f.nnz * sizeof(double), cudaMemcpyDeviceToHost));
  }
  pf.n = f.n;
  pf.W = std::max(f.maxrow, 1);
  std::map<double, int> ids;
  std::vector<double> vtab(1, 0.0);
  std::vector<uint32_t> ell((size_t)pf.W * f.n);
  for (int64_t i = 0; i < f.n; i++) {
    int k = 0;
    for (int32_t p = rp[i]; p < rp[i + 1]; p++, k++) {
      double av = vals[p] < 0 ? -vals[p] : vals[p];
      auto it = ids.find(av);
      int id;
      if (it == ids.end()) { id = (int)vtab.size(); ids[av] = id; vtab.push_back(av); }
      else id = it->second;
      if (id > (int)ELL_VID_MASK) return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "too many distinct matrix elements for the packed factor");
      ell[(size_t)k * f.n + i] = (uint32_t)cols[p] | ((uint32_t)id << ELL_COL_BITS) | (vals[p] < 0 ? 0x80000000u : 0u);
    }
    for (; k < pf.W; k++) ell[(size_t)k * f.n + i] = (uint32_t)i;          // padding: value id 0 = 0.0
  }
  pf.nvals = (int)vtab.size();
  CK(cudaMalloc(&pf.d_ell, ell.size() * sizeof(uint32_t)));
  CK(cudaMalloc(&pf.d_vtab, vtab.size() * sizeof(double)));
  CK(cudaMemcpy(pf.d_ell, ell.data(), ell.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(pf.d_vtab, vtab.data(), vtab.size() * sizeof(double), cudaMemcpyHostToDevice));
  return EDGPU_OK;
}

static size_t col_smem(int G, int W, int64_t chunk, bool direct) {
  size_t b = (size_t)G * chunk * 8 + (direct ? (size_t)chunk * 8 : 0) + TILED_MAX_VALS * 8 + (size_t)G * (1 + EDGPU_MAX_ORB) * 8 +
             (size_t)W * chunk * 4 + (direct ? (size_t)chunk * 4 : 0);
  return (b + 15) & ~(size_t)15;
}
static size_t row_smem(int R, int W, int64_t chunk) {
  size_t b = (size_t)chunk * R * 8 + TILED_MAX_VALS * 8 + (size_t)W * chunk * 4;
  return (b + 15) & ~(size_t)15;
}

bool tiled_supported(const edgpu_ctx *c) {
  if (!c->hstatus || c->dp.jhflag) return false;
  if (c->dimup >= (1 << ELL_COL_BITS) || c->dimdw >= (1 << ELL_COL_BITS)) return false;
  if (c->up.maxrow > TILED_MAX_W || c->dw.maxrow > TILED_MAX_W) return false;
  return true;
}

int tiled_plan_free(edgpu_ctx *c) {
  if (!c->plan) return EDGPU_OK;
  for (int k = 0; k < 2; k++) { cudaFree(c->plan->pf[k].d_ell); cudaFree(c->plan->pf[k].d_vtab); }
  delete c->plan;
  c->plan = nullptr;
  return EDGPU_OK;
}

int tiled_plan_build(edgpu_ctx *c) {
  if (c->plan) return EDGPU_OK;
  TiledPlan *p = new TiledPlan();
  c->plan = p;
  int rc = pack_factor(c, c->up, p->pf[0]);
  if (!rc) rc = pack_factor(c, c->dw, p->pf[1]);
  if (!rc && (p->pf[0].nvals > TILED_MAX_VALS || p->pf[1].nvals > TILED_MAX_VALS))
    rc = edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "packed factor value table too large");
  if (rc) { tiled_plan_free(c); return rc; }
  const bool direct = c->d_diag == nullptr;
  // pass-1 style chunking for both factors (the dw one is used on the transposed shard)
  for (int k = 0; k < 2; k++) {
    const int64_t n = p->pf[k].n;
    const int W = p->pf[k].W;
    const bool dg = direct && k == 0;
    int64_t want = c->opt_col_h > 0 ? c->opt_col_h : 0;       // option col_h: number of row chunks
    int best_nch = 0, best_G = 1;
    for (int nch = (want > 0 ? (int)want : 1); nch <= 4096; nch++) {
      int64_t chunk = ((n + nch - 1) / nch + 1) & ~(int64_t)1;
      int G = 0;
      for (int g : {8, 4, 2, 1}) if (col_smem(g, W, chunk, dg) <= SMEM_BUDGET) { G = g; break; }
      if (G >= 2 || (G == 1 && want > 0)) { best_nch = nch; best_G = G; break; }
      if (G == 1 && best_nch == 0) { best_nch = nch; best_G = 1; }   // remember, but prefer G >= 2
      if (want > 0) break;
    }
    if (best_nch == 0) { tiled_plan_free(c); return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "no column-kernel tiling fits shared memory"); }
    p->a_nchunks[k] = best_nch;
    p->a_G[k] = best_G;
    p->a_chunk[k] = ((n + best_nch - 1) / best_nch + 1) & ~(int64_t)1;
    p->a_nchunks[k] = (int)((n + p->a_chunk[k] - 1) / p->a_chunk[k]);      // even rounding can drop a chunk
    p->a_smem[k] = col_smem(best_G, W, p->a_chunk[k], dg);
  }
  // pass 2: R rows x chunk of dw columns
  {
    const int R = c->opt_tile_rows == 8 ? 8 : 16;
    const int64_t n = p->pf[1].n;
    const int W = p->pf[1].W;
    int nch = c->opt_tile_h > 0 ? (int)c->opt_tile_h : 1;
    for (; nch <= 1 << 16; nch++) {
      int64_t chunk = (n + nch - 1) / nch;
      if (row_smem(R, W, chunk) <= SMEM_BUDGET) break;
      if (c->opt_tile_h > 0) { tiled_plan_free(c); return edgpu_set_err(EDGPU_ERR_INVALID, "tile_h does not fit shared memory"); }
    }
    p->b_R = R;
    p->b_nchunks = nch;
    p->b_chunk = (n + nch - 1) / nch;
    p->b_nchunks = (int)((n + p->b_chunk - 1) / p->b_chunk);
    p->b_smem = row_smem(R, W, p->b_chunk);
  }
  if (!c->tiled_attrs_set) {                                      // per-device function attribute: once per context
    const int mx = (int)SMEM_BUDGET;
#define SETA(k) CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, mx))
    SETA((k_tile_col<1, 0>)); SETA((k_tile_col<2, 0>)); SETA((k_tile_col<4, 0>)); SETA((k_tile_col<8, 0>));
    SETA((k_tile_col<1, 1>)); SETA((k_tile_col<2, 1>)); SETA((k_tile_col<4, 1>)); SETA((k_tile_col<8, 1>));
    SETA((k_tile_col<1, 2>)); SETA((k_tile_col<2, 2>)); SETA((k_tile_col<4, 2>)); SETA((k_tile_col<8, 2>));
    SETA((k_tile_row<8>)); SETA((k_tile_row<16>));
#undef SETA
    c->tiled_attrs_set = true;
  }
  return EDGPU_OK;
}

template <int DIAG>
static void launch_col(int G, dim3 grid, size_t smem, cudaStream_t st, const ColArgs &a) {
  switch (G) {
    case 8: k_tile_col<8, DIAG><<<grid, TILED_THREADS, smem, st>>>(a); break;
    case 4: k_tile_col<4, DIAG><<<grid, TILED_THREADS, smem, st>>>(a); break;
    case 2: k_tile_col<2, DIAG><<<grid, TILED_THREADS, smem, st>>>(a); break;
    default: k_tile_col<1, DIAG><<<grid, TILED_THREADS, smem, st>>>(a); break;
  }
}

// y = [Hd o x +] F_k x on a matrix whose contiguous dimension is factor k's index
int tiled_apply_col(edgpu_ctx *c, int k, bool with_diag, const double *d_x, double *d_y, int64_t ncols, int64_t coloff) {
  TRY(tiled_plan_build(c));
  TiledPlan *p = c->plan;
  ColArgs a{};
  a.x = d_x; a.y = d_y; a.n = p->pf[k].n; a.ncols = ncols; a.coloff = coloff;
  a.ell = p->pf[k].d_ell; a.vtab = p->pf[k].d_vtab; a.W = p->pf[k].W; a.nvals = p->pf[k].nvals;
  a.chunk = p->a_chunk[k]; a.nchunks = p->a_nchunks[k];
  const int64_t ngroups = (ncols + p->a_G[k] - 1) / p->a_G[k];
  int cpc = std::max(1, (2 * c->sm_count) / a.nchunks);
  if (cpc > ngroups) cpc = (int)ngroups;
  a.cpc = cpc;
  a.diag = c->d_diag;
  a.dfac_c = c->up.d_dfac; a.dfac_s = c->dw.d_dfac; a.map_c = c->up.d_map; a.map_s = c->dw.d_map;
  a.norb = c->dp.norb; a.ust = c->dp.ust;
  for (int i = 0; i < EDGPU_MAX_ORB; i++) a.uloc[i] = c->dp.uloc[i];
  dim3 grid((unsigned)(a.nchunks * cpc));
  const int diag = !with_diag ? 0 : (c->d_diag ? 1 : 2);
  if (diag == 0) launch_col<0>(p->a_G[k], grid, p->a_smem[k], c->stream, a);
  else if (diag == 1) launch_col<1>(p->a_G[k], grid, p->a_smem[k], c->stream, a);
  else launch_col<2>(p->a_G[k], grid, p->a_smem[k], c->stream, a);
  CKL(c);
  return EDGPU_OK;
}

// y += x Hdw^T on the full local matrix (all i_dw columns local)
static int tiled_apply_row(edgpu_ctx *c, const double *d_x, double *d_y) {
  TiledPlan *p = c->plan;
  RowArgs a{};
  a.x = d_x; a.y = d_y; a.n = c->dimup; a.nf = c->dimdw;
  a.ell = p->pf[1].d_ell; a.vtab = p->pf[1].d_vtab; a.W = p->pf[1].W; a.nvals = p->pf[1].nvals;
  a.chunk = p->b_chunk; a.nchunks = p->b_nchunks;
  const int64_t ntiles = (a.n + p->b_R - 1) / p->b_R;
  int cpc = std::max(1, (2 * c->sm_count) / a.nchunks);
  if (cpc > ntiles) cpc = (int)ntiles;
  a.cpc = cpc;
  dim3 grid((unsigned)(a.nchunks * cpc));
  if (p->b_R == 8) k_tile_row<8><<<grid, TILED_THREADS, p->b_smem, c->stream>>>(a);
  else k_tile_row<16><<<grid, TILED_THREADS, p->b_smem, c->stream>>>(a);
  CKL(c);
  return EDGPU_OK;
}

int tiled_apply_local(edgpu_ctx *c, const double *d_x, double *d_y) {
  TRY(tiled_plan_build(c));
  prof_mark(c, "k_tile_col");
  TRY(tiled_apply_col(c, 0, true, d_x, d_y, c->qdw, c->coloff));
  prof_mark(c, "k_tile_row");
  TRY(tiled_apply_row(c, d_x, d_y));
  return EDGPU_OK;
}
