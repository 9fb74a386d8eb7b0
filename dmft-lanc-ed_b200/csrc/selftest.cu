// selftest.cu -- test instrumentation, built into its OWN library (libedgpu_selftest.so, linked against
// libedgpu.so; include/edgpu_selftest.h): HOST evaluation of the __host__ __device__ integer/bit logic in
// hd_funcs.h and of the row-kernel plan, so that the CPU-only test-suite can check the exact code the kernels run
// (rank formula, factor rows, diagonal order, chunk / record / list arithmetic) against the oracle without a GPU,
// plus the one-device emulation of the sharded path.  The product library exports none of this.
#include <string.h>

#include <vector>

#include "../../include/edgpu_selftest.h"
#include "engine.h"

static void host_binom(uint32_t *b) {
  memset(b, 0, sizeof(uint32_t) * EDGPU_BINOM_LD * EDGPU_BINOM_LD);
  for (int n = 0; n < EDGPU_BINOM_LD; n++) {
    b[n * EDGPU_BINOM_LD] = 1;
    for (int k = 1; k <= n; k++)
      b[n * EDGPU_BINOM_LD + k] = b[(n - 1) * EDGPU_BINOM_LD + k - 1] + (k <= n - 1 ? b[(n - 1) * EDGPU_BINOM_LD + k] : 0);
  }
}
static DevParams make_dp(const edgpu_params *p) {
  DevParams d;
  memset(&d, 0, sizeof(d));
  d.norb = p->norb; d.nbath = p->nbath; d.ns = (p->nbath + 1) * p->norb; d.hfmode = p->hfmode; d.nspin = p->nspin;
  d.bath_type = 0; d.nfoo = p->norb;                               // the host self-tests cover the normal bath
  d.jhflag = (p->norb > 1 && (p->jx != 0.0 || p->jp != 0.0));
  for (int i = 0; i < EDGPU_MAX_ORB; i++) d.uloc[i] = (i < p->norb) ? p->uloc[i] : 0.0;
  d.ust = p->ust; d.jh = p->jh; d.jx = p->jx; d.jp = p->jp; d.xmu = p->xmu;
  const int nsn = p->nspin, sl = p->nspin - 1;
  for (int io = 0; io < p->norb; io++)
    for (int jo = 0; jo < p->norb; jo++) {
      d.hloc_up[io * EDGPU_MAX_ORB + jo] = p->imphloc ? p->imphloc[0 + nsn * (0 + nsn * (io + p->norb * jo))] : 0.0;
      d.hloc_dw[io * EDGPU_MAX_ORB + jo] = p->imphloc ? p->imphloc[sl + nsn * (sl + nsn * (io + p->norb * jo))] : 0.0;
    }
  for (int io = 0; io < p->norb; io++)
    for (int kp = 0; kp < p->nbath; kp++) {
      d.be_up[io * p->nbath + kp] = p->bath_e[0 + nsn * (io + p->norb * kp)];
      d.be_dw[io * p->nbath + kp] = p->bath_e[sl + nsn * (io + p->norb * kp)];
      d.bv_up[io * p->nbath + kp] = p->bath_v[0 + nsn * (io + p->norb * kp)];
      d.bv_dw[io * p->nbath + kp] = p->bath_v[sl + nsn * (io + p->norb * kp)];
    }
  return d;
}

extern "C" int64_t edgpu_selftest_map(int ns, int n, int32_t *map) {
  uint32_t b[EDGPU_BINOM_LD * EDGPU_BINOM_LD];
  host_binom(b);
  int64_t dim = b[ns * EDGPU_BINOM_LD + n];
  if (!map) return dim;
  for (uint64_t s = 0; s < (1ull << ns); s++)
    if (hd_popc((uint32_t)s) == n) map[hd_rank((uint32_t)s, b)] = (int32_t)s;
  return dim;
}

extern "C" int64_t edgpu_selftest_factor(const edgpu_params *p, int spin, int npart, int64_t *rowptr,
                                         int64_t *cols, double *vals) {
  DevParams d = make_dp(p);
  int64_t n = edgpu_selftest_map(d.ns, npart, nullptr);
  std::vector<int32_t> map((size_t)n);
  edgpu_selftest_map(d.ns, npart, map.data());
  int64_t nnz = 0;
  int32_t c[EDGPU_MAX_ROW_NNZ]; double v[EDGPU_MAX_ROW_NNZ];
  for (int64_t i = 0; i < n; i++) {
    int m = hd_factor_row(d, spin, map.data(), n, (uint32_t)map[i], c, v);
    if (rowptr) {
      rowptr[i] = nnz;
      for (int k = 0; k < m; k++) { cols[nnz + k] = c[k]; vals[nnz + k] = v[k]; }
    }
    nnz += m;
  }
  if (rowptr) rowptr[n] = nnz;
  return nnz;
}

extern "C" double edgpu_selftest_diag(const edgpu_params *p, uint32_t mup, uint32_t mdw, int factorised) {
  DevParams d = make_dp(p);
  if (!factorised) return hd_diag_element(d, mup, mdw);
  return hd_diag_factor(d, 0, mup) + hd_diag_factor(d, 1, mdw) + hd_diag_cross(d, mup, mdw);
}

extern "C" int edgpu_selftest_nonlocal_row(const edgpu_params *p, uint32_t mup, uint32_t mdw, uint32_t *cup,
                                           uint32_t *cdw, double *val) {
  DevParams d = make_dp(p);
  return hd_nonlocal_row(d, mup, mdw, cup, cdw, val);
}

// max |a - b| and max |a| of two device vectors (full-size comparisons without a host round trip)
__global__ void k_maxabsdiff(const double *__restrict__ a, const double *__restrict__ b, int64_t n, unsigned long long *out) {
  double md = 0.0, ma = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    md = fmax(md, fabs(a[i] - b[i]));
    ma = fmax(ma, fabs(a[i]));
  }
  if (md != md) md = 1e300;                                        // NaN counts as a mismatch
  for (int o = 16; o > 0; o >>= 1) { md = fmax(md, __shfl_xor_sync(0xffffffffu, md, o)); ma = fmax(ma, __shfl_xor_sync(0xffffffffu, ma, o)); }
  if ((threadIdx.x & 31) == 0) {                                   // non-negative doubles order like their bit patterns
    atomicMax(&out[0], (unsigned long long)__double_as_longlong(md));
    atomicMax(&out[1], (unsigned long long)__double_as_longlong(ma));
  }
}
extern "C" int edgpu_selftest_dev_maxabsdiff(const double *d_a, const double *d_b, int64_t n, double *maxdiff, double *maxabs) {
  unsigned long long *d_out = nullptr, h[2] = {0, 0};
  if (cudaMalloc(&d_out, 16) != cudaSuccess) return 1;
  cudaMemset(d_out, 0, 16);
  k_maxabsdiff<<<1184, 256>>>(d_a, d_b, n, d_out);
  if (cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaFree(d_out); return 1; }
  cudaFree(d_out);
  memcpy(maxdiff, &h[0], 8);
  memcpy(maxabs, &h[1], 8);
  return 0;
}

// Sharded fast path on ONE device: `nranks` contexts stand in for the ranks (no NCCL, no IPC: every "rank" gets
// a plain allocation as its halo slab and the others' addresses by hand; no arrival flags -- the pushes of all ranks
// run first, then every rank's H*v).  Exercises exactly the kernels and plans of the multi-GPU path -- whole / cut
// low groups, group records, halo push + sum kernels, the source lists of the column pass -- so that a single-GPU
// box can check them for any rank count.  x, y: full host vectors.
extern "C" int edgpu_selftest_sharded_hxv(const edgpu_params *p, int nup, int ndw, int nranks, int64_t srow_lr,
                                          int64_t srow_t, int64_t col_cluster, const double *x, double *y) {
  if (nranks < 1 || nranks > EDGPU_MAXP) return edgpu_set_err(EDGPU_ERR_INVALID, "selftest: 1 <= nranks <= 8");
  std::vector<edgpu_ctx *> cs((size_t)nranks, nullptr);
  std::vector<double *> dx((size_t)nranks, nullptr), dy((size_t)nranks, nullptr);
  std::vector<char *> slab((size_t)nranks, nullptr);
  int rc = EDGPU_OK;
  auto cleanup = [&]() {
    for (int r = 0; r < nranks; r++) {
      if (dx[r]) cudaFree(dx[r]);
      if (dy[r]) cudaFree(dy[r]);
      if (slab[r]) cudaFree(slab[r]);
      if (cs[r]) {
        cs[r]->sym_slab = nullptr; cs[r]->sym_ok = false; cs[r]->halo_emul = false;
        for (int q = 0; q < 64; q++) cs[r]->sym_peer[q] = nullptr;
        cs[r]->rank = 0; cs[r]->nranks = 1;
        edgpu_destroy(cs[r]);
      }
    }
  };
  for (int r = 0; r < nranks && !rc; r++) {
    rc = edgpu_create(p, -1, &cs[r]);
    if (rc) break;
    cs[r]->rank = r; cs[r]->nranks = nranks;                       // no communicator: peers are plain pointers here
    edgpu_set_option(cs[r], "srow_lr", srow_lr);
    edgpu_set_option(cs[r], "srow_t", srow_t);
    edgpu_set_option(cs[r], "col_cluster", col_cluster);
    edgpu_set_option(cs[r], "no_peer", 1);                         // no collective slab setup without a communicator
    int isec = 0;
    rc = edgpu_get_sector(cs[r], nup, ndw, &isec);
    if (!rc) rc = edgpu_build_hv_sector(cs[r], isec);
    if (rc) break;
    const size_t nb = ((size_t)cs[r]->nloc + 2) * sizeof(double);
    if (cudaMalloc(&dx[r], nb) != cudaSuccess || cudaMalloc(&dy[r], nb) != cudaSuccess) { rc = edgpu_set_err(EDGPU_ERR_CUDA, "selftest: cudaMalloc"); break; }
    cudaMemset(dy[r], 0xff, nb);                                   // NaN pattern: every element must be written
    if (cudaMemcpy(dx[r], x + cs[r]->coloff * cs[r]->dimup, (size_t)cs[r]->nloc * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
      rc = edgpu_set_err(EDGPU_ERR_CUDA, "selftest: upload");
      break;
    }
  }
  if (!rc) cudaDeviceSynchronize();
  if (nranks > 1) {
    size_t hb0 = 0;
    for (int r = 0; r < nranks && !rc; r++) {
      size_t hb = 0;
      if (!fast_supported_local(cs[r])) { rc = edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "selftest: fast path does not cover this sector"); break; }
      rc = fast_halo_bytes(cs[r], &hb);
      if (!rc && r > 0 && hb != hb0) rc = edgpu_set_err(EDGPU_ERR_INVALID, "selftest: halo slab size differs between ranks");
      hb0 = hb;
      if (!rc && hb && cudaMalloc(&slab[r], hb) != cudaSuccess) rc = edgpu_set_err(EDGPU_ERR_CUDA, "selftest: cudaMalloc of the halo slab");
      if (!rc && hb) cudaMemset(slab[r], 0xff, hb);                // NaN pattern: a slot that nobody stores shows up
    }
    for (int r = 0; r < nranks && !rc && hb0; r++) {
      cs[r]->sym_slab = slab[r]; cs[r]->sym_bytes = hb0; cs[r]->sym_ok = true; cs[r]->halo_emul = true;
      for (int q = 0; q < nranks; q++) cs[r]->sym_peer[q] = slab[q];
      cs[r]->halo_epoch = 1;
    }
    for (int r = 0; r < nranks && !rc && hb0; r++) {
      rc = fast_halo_push(cs[r], dx[r], cs[r]->stream, 64, -1);
      if (!rc && cudaDeviceSynchronize() != cudaSuccess) rc = edgpu_set_err(EDGPU_ERR_CUDA, "selftest: push: %s", cudaGetErrorString(cudaGetLastError()));
    }
  }
  for (int r = 0; r < nranks && !rc; r++) {
    if (!fast_supported_local(cs[r])) { rc = edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "selftest: fast path does not cover this sector"); break; }
    rc = fast_apply_local(cs[r], dx[r], dy[r], nullptr, nullptr);
    if (!rc && cudaDeviceSynchronize() != cudaSuccess) rc = edgpu_set_err(EDGPU_ERR_CUDA, "selftest: kernels: %s", cudaGetErrorString(cudaGetLastError()));
  }
  for (int r = 0; r < nranks && !rc; r++)
    if (cudaMemcpy(y + cs[r]->coloff * cs[r]->dimup, dy[r], (size_t)cs[r]->nloc * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess)
      rc = edgpu_set_err(EDGPU_ERR_CUDA, "selftest: download");
  cleanup();
  return rc;
}

// HOST ONLY (no device): the plan of the structured row kernel for `rank` of `nranks` -- Lin table, chunk table,
// group records and the source lists of the hops the row kernel leaves out -- exactly as build_Hv_sector computes it,
// so that the CPU-only tests can check the chunking / sharding logic.  info[8] = {ok, LR, T, nhigh, nchunks, nrecs,
// list entries, listed columns}; arrays may be NULL; capacities in entries.  recs: 20 int32 per record (lb, N, hx,
// par, pc[16]).  lptr has qdw+1 entries, lflag qdw.  Returns 0, or 1 when a capacity was too small.
extern "C" int edgpu_selftest_srow_plan(const edgpu_params *p, int ndw, int nranks, int rank, int64_t lr, int64_t tbits_opt,
                                        int32_t *info, int32_t *jhi, int cap_jhi, int32_t *chunks, int cap_chunks,
                                        int32_t *recs, int cap_recs, int32_t *lptr, int32_t *lflag, int cap_cols,
                                        int32_t *lown, int32_t *lcol, double *lamp, int cap_e) {
  DevParams d = make_dp(p);
  const int64_t n = edgpu_selftest_map(d.ns, ndw, nullptr);
  SRowHostPlan hp;
  for (int k = 0; k < 8; k++) info[k] = 0;
  const int rc = srow_plan_host(d.ns, ndw, n, nranks, rank, (int)lr, (int)tbits_opt, hp);
  if (rc <= 0) { info[0] = rc; return 0; }
  std::vector<int32_t> map((size_t)n), rp((size_t)n + 1), cc;
  std::vector<double> vv;
  edgpu_selftest_map(d.ns, ndw, map.data());
  int32_t c[EDGPU_MAX_ROW_NNZ]; double v[EDGPU_MAX_ROW_NNZ];
  for (int64_t i = 0; i < n; i++) {
    rp[(size_t)i] = (int32_t)cc.size();
    const int m = hd_factor_row(d, 1, map.data(), n, (uint32_t)map[(size_t)i], c, v);
    for (int k = 0; k < m; k++) { cc.push_back(c[k]); vv.push_back(v[k]); }
  }
  rp[(size_t)n] = (int32_t)cc.size();
  srow_lists_host(hp, rank, n, map.data(), rp.data(), cc.data(), vv.data());
  info[0] = 1; info[1] = hp.LR; info[2] = hp.T; info[3] = hp.nhigh; info[4] = (int32_t)hp.chunks.size();
  info[5] = (int32_t)hp.recs.size(); info[6] = (int32_t)hp.lown.size(); info[7] = (int32_t)hp.zcols.size();
  int small = 0;
  if (jhi) { if ((int)hp.jhi.size() > cap_jhi) small = 1; else memcpy(jhi, hp.jhi.data(), hp.jhi.size() * sizeof(int32_t)); }
  if (chunks) {
    if ((int)hp.chunks.size() > cap_chunks) small = 1;
    else for (size_t k = 0; k < hp.chunks.size(); k++) { chunks[4 * k] = hp.chunks[k].x; chunks[4 * k + 1] = hp.chunks[k].y; chunks[4 * k + 2] = hp.chunks[k].z; chunks[4 * k + 3] = hp.chunks[k].w; }
  }
  if (recs) { if ((int)hp.recs.size() > cap_recs) small = 1; else if (!hp.recs.empty()) memcpy(recs, hp.recs.data(), hp.recs.size() * sizeof(SRowRec)); }
  if (lptr) {
    if ((int)hp.lflag.size() > cap_cols) small = 1;
    else { for (size_t k = 0; k < hp.lptr.size(); k++) lptr[k] = hp.lptr[k]; for (size_t k = 0; k < hp.lflag.size(); k++) lflag[k] = hp.lflag[k]; }
  }
  if (lown) {
    if ((int)hp.lown.size() > cap_e) small = 1;
    else for (size_t k = 0; k < hp.lown.size(); k++) { lown[k] = hp.lown[k]; lcol[k] = hp.lcol[k]; lamp[k] = hp.lamp[k]; }
  }
  return small;
}

// HOST ONLY: the halo tables of the push model for EVERY rank of `nranks`, cross-checked against each other: every
// remote list entry (rank q, slot s, owner p, source column) must be stored by exactly one triple of rank p, in the
// window of its target column, and no triple may exist without an entry.  info[6] = {ok, list entries, remote
// entries, triples, largest slot count, windows used}.
extern "C" int edgpu_selftest_halo_tables(const edgpu_params *p, int ndw, int nranks, int64_t lr, int64_t tbits_opt, int K, int32_t *info) {
  for (int k = 0; k < 6; k++) info[k] = 0;
  DevParams d = make_dp(p);
  const int64_t n = edgpu_selftest_map(d.ns, ndw, nullptr);
  std::vector<int32_t> map((size_t)n), rp((size_t)n + 1), cc;
  std::vector<double> vv;
  edgpu_selftest_map(d.ns, ndw, map.data());
  int32_t cb[EDGPU_MAX_ROW_NNZ]; double vb[EDGPU_MAX_ROW_NNZ];
  for (int64_t i = 0; i < n; i++) {
    rp[(size_t)i] = (int32_t)cc.size();
    const int m = hd_factor_row(d, 1, map.data(), n, (uint32_t)map[(size_t)i], cb, vb);
    for (int k = 0; k < m; k++) { cc.push_back(cb[k]); vv.push_back(vb[k]); }
  }
  rp[(size_t)n] = (int32_t)cc.size();
  struct Rank { SRowHostPlan hp; std::vector<int> lcol2, pdst, pslot, psrc; int pwin[EDGPU_MAX_WINDOWS + 1]; int nslot, maxslot; };
  std::vector<Rank> R((size_t)nranks);
  for (int r = 0; r < nranks; r++) {
    if (srow_plan_host(d.ns, ndw, n, nranks, r, (int)lr, (int)tbits_opt, R[(size_t)r].hp) != 1) return 0;   // info[0] = 0
    srow_lists_host(R[(size_t)r].hp, r, n, map.data(), rp.data(), cc.data(), vv.data());
    if (halo_tables_host(d.ns, ndw, n, nranks, r, (int)lr, (int)tbits_opt, R[(size_t)r].hp, map.data(), rp.data(), cc.data(), vv.data(), K,
                         R[(size_t)r].lcol2, R[(size_t)r].pdst, R[(size_t)r].pslot, R[(size_t)r].psrc, R[(size_t)r].pwin, &R[(size_t)r].nslot,
                         &R[(size_t)r].maxslot)) return 0;
  }
  long entries = 0, remote = 0, triples = 0;
  bool ok = true;
  int maxslot = 0;
  for (int q = 0; q < nranks && ok; q++) {
    const Rank &Q = R[(size_t)q];
    const int64_t qq = (int64_t)Q.hp.lptr.size() - 1;
    if (Q.maxslot != R[0].maxslot) ok = false;                     // the slab layout must be the same everywhere
    maxslot = std::max(maxslot, Q.nslot);
    std::vector<char> hit((size_t)std::max(Q.nslot, 1), 0);
    int slot = 0;
    for (int64_t t = 0; t < qq && ok; t++) {
      int w = 0;
      while (w + 1 < K && t >= qq * (w + 1) / K) w++;
      for (int e = Q.hp.lptr[(size_t)t]; e < Q.hp.lptr[(size_t)t + 1] && ok; e++) {
        entries++;
        const int own = Q.hp.lown[(size_t)e];
        if (own == q) { if (Q.lcol2[(size_t)e] != Q.hp.lcol[(size_t)e]) ok = false; continue; }
        remote++;
        if (Q.lcol2[(size_t)e] != slot) ok = false;
        // the owner's triple for (q, slot): in window w, with the entry's source column
        const Rank &O = R[(size_t)own];
        int found = 0;
        for (int k = O.pwin[w]; k < O.pwin[w + 1]; k++)
          if (O.pdst[(size_t)k] == q && O.pslot[(size_t)k] == slot) { found++; if (O.psrc[(size_t)k] != Q.hp.lcol[(size_t)e]) ok = false; }
        if (found != 1) ok = false;
        hit[(size_t)slot] = 1;
        slot++;
      }
    }
    if (slot != Q.nslot) ok = false;
    triples += (long)Q.pdst.size();
    if (Q.pwin[0] != 0 || Q.pwin[K] != (int)Q.pdst.size()) ok = false;
    for (int w = 0; w < K; w++) {
      if (Q.pwin[w] > Q.pwin[w + 1]) ok = false;
      for (int k = Q.pwin[w] + 1; k < Q.pwin[w + 1]; k++) if (Q.psrc[(size_t)k] < Q.psrc[(size_t)k - 1]) ok = false;   // sorted by source column
    }
  }
  if (triples != remote || maxslot != R[0].maxslot) ok = false;
  info[0] = ok ? 1 : 0; info[1] = (int32_t)entries; info[2] = (int32_t)remote; info[3] = (int32_t)triples; info[4] = maxslot; info[5] = K;
  return 0;
}

// HOST ONLY: the halo tables of ONE rank as build_Hv_sector computes them (the world_size-2 gloo test drives a real
// inter-process exchange with them).  info[4] = {slots of this rank, largest slot count of any rank, triples, list
// entries}; lcol2[cap_e]; pdst / pslot / psrc[cap_p]; pwin[K + 1].  Returns 0, 1 when a capacity was too small, -1 when
// the structured row kernel does not apply.
extern "C" int edgpu_selftest_halo_rank(const edgpu_params *p, int ndw, int nranks, int rank, int64_t lr, int64_t tbits_opt, int K,
                                        int32_t *info, int32_t *lcol2, int cap_e, int32_t *pdst, int32_t *pslot, int32_t *psrc,
                                        int cap_p, int32_t *pwin) {
  DevParams d = make_dp(p);
  const int64_t n = edgpu_selftest_map(d.ns, ndw, nullptr);
  std::vector<int32_t> map((size_t)n), rp((size_t)n + 1), cc;
  std::vector<double> vv;
  edgpu_selftest_map(d.ns, ndw, map.data());
  int32_t cb[EDGPU_MAX_ROW_NNZ]; double vb[EDGPU_MAX_ROW_NNZ];
  for (int64_t i = 0; i < n; i++) {
    rp[(size_t)i] = (int32_t)cc.size();
    const int m = hd_factor_row(d, 1, map.data(), n, (uint32_t)map[(size_t)i], cb, vb);
    for (int k = 0; k < m; k++) { cc.push_back(cb[k]); vv.push_back(vb[k]); }
  }
  rp[(size_t)n] = (int32_t)cc.size();
  SRowHostPlan hp;
  if (srow_plan_host(d.ns, ndw, n, nranks, rank, (int)lr, (int)tbits_opt, hp) != 1) return -1;
  srow_lists_host(hp, rank, n, map.data(), rp.data(), cc.data(), vv.data());
  std::vector<int> l2, pd, ps, pr;
  int pw[EDGPU_MAX_WINDOWS + 1], nslot = 0, maxslot = 0;
  if (K < 1 || K > EDGPU_MAX_WINDOWS) return -1;
  if (halo_tables_host(d.ns, ndw, n, nranks, rank, (int)lr, (int)tbits_opt, hp, map.data(), rp.data(), cc.data(), vv.data(), K, l2, pd, ps, pr,
                       pw, &nslot, &maxslot)) return -1;
  info[0] = nslot; info[1] = maxslot; info[2] = (int32_t)pd.size(); info[3] = (int32_t)l2.size();
  if ((int)l2.size() > cap_e || (int)pd.size() > cap_p) return 1;
  for (size_t k = 0; k < l2.size(); k++) lcol2[k] = l2[k];
  for (size_t k = 0; k < pd.size(); k++) { pdst[k] = pd[k]; pslot[k] = ps[k]; psrc[k] = pr[k]; }
  for (int w = 0; w <= K; w++) pwin[w] = pw[w];
  return 0;
}

// HOST ONLY: factor f of an ed_total_ud = F sector exactly as build_Hv_sector computes it (f < Norb: up word of orbital
// f+1 with n electrons, else the dw word of orbital f+1-Norb).  Returns the dimension; arrays may be NULL (count only);
// *nnz = number of entries.
extern "C" int64_t edgpu_selftest_orbs_factor(const edgpu_params *p, int f, int n, int32_t *map, int64_t *nnz, int64_t *rowptr,
                                              int64_t *cols, double *vals) {
  DevParams d = make_dp(p);
  std::vector<int32_t> m, rp, cc;
  std::vector<double> vv;
  orbs_factor_host(d, f, n, m, rp, cc, vv);
  if (nnz) *nnz = (int64_t)cc.size();
  if (map) for (size_t k = 0; k < m.size(); k++) map[k] = m[k];
  if (rowptr) for (size_t k = 0; k < rp.size(); k++) rowptr[k] = rp[k];
  if (cols) for (size_t k = 0; k < cc.size(); k++) cols[k] = cc[k];
  if (vals) for (size_t k = 0; k < vv.size(); k++) vals[k] = vv[k];
  return (int64_t)m.size();
}
