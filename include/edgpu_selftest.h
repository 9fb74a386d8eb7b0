/*
 * edgpu_selftest.h -- test instrumentation exported by libedgpu_selftest.so (a separate library linked against
 * libedgpu.so; the product library exports none of it): HOST evaluation of the
 * __host__ __device__ bit logic shared with the kernels (dmft-lanc-ed_b200/csrc/hd_funcs.h).
 * Not part of the drop-in boundary; no product entry point calls these and they do not compute
 * H*v.  They let the CPU-only test-suite compare the kernels' index/sign/diagonal code with the
 * oracle (ED_SETUP.f90:745-831,1042-1059; ED_HAMILTONIAN/stored/H_local.f90, H_up.f90,
 * H_non_local.f90) before any GPU time is spent.
 */
#ifndef EDGPU_SELFTEST_H
#define EDGPU_SELFTEST_H
#include <stdint.h>
#include "edgpu.h"
#ifdef __cplusplus
extern "C" {
#endif
/* sector map by the closed-form rank (build_sector); map==NULL returns the dimension */
int64_t edgpu_selftest_map(int ns, int n, int32_t *map);
/* CSR of spH0ups(1) (spin=0) / spH0dws(1) (spin=1); rowptr==NULL returns nnz */
int64_t edgpu_selftest_factor(const edgpu_params *p, int spin, int npart, int64_t *rowptr, int64_t *cols,
                              double *vals);
/* diagonal element: factorised=0 reference summation order, 1 the factorised form of direct mode */
double edgpu_selftest_diag(const edgpu_params *p, uint32_t mup, uint32_t mdw, int factorised);
/* one row of spH0nd: returns the entry count, outputs column words and values */
int edgpu_selftest_nonlocal_row(const edgpu_params *p, uint32_t mup, uint32_t mdw, uint32_t *cup, uint32_t *cdw,
                                double *val);
/* HOST: plan of the structured row kernel for `rank` of `nranks` (Lin table, chunk table, group records, the source
 * lists of the hops the row kernel leaves out on a sharded vector), as build_Hv_sector computes it.  tbits_opt = t + 1
 * forces chunks of 2^t low groups (0 = automatic).  info[8] = {ok, LR, T, nhigh, nchunks, nrecs, list entries, listed
 * columns}; arrays may be NULL; recs: 20 int32 per record (lb, N, hx, par, pc[16]); lptr: qdw+1 entries; lflag: qdw
 * (bit 0: column not written by the row kernel, bit 1: column has entries); entry = (owner rank, column inside the
 * owner's shard, amplitude). */
int edgpu_selftest_srow_plan(const edgpu_params *p, int ndw, int nranks, int rank, int64_t lr, int64_t tbits_opt,
                             int32_t *info, int32_t *jhi, int cap_jhi, int32_t *chunks, int cap_chunks,
                             int32_t *recs, int cap_recs, int32_t *lptr, int32_t *lflag, int cap_cols,
                             int32_t *lown, int32_t *lcol, double *lamp, int cap_e);
/* HOST ONLY: the halo tables of the sharded fast path (push model) of EVERY rank, cross-checked: each remote list entry
 * (rank q, slot s) is stored by exactly one (peer, slot, source column) triple of its owner, in the column window of
 * its target; no stray triples; same slab layout on every rank.  info[6] = {ok, list entries, remote entries, triples,
 * largest slot count, windows}. */
int edgpu_selftest_halo_tables(const edgpu_params *p, int ndw, int nranks, int64_t lr, int64_t tbits_opt, int K, int32_t *info);
/* HOST ONLY: factor f of an ed_total_ud = F sector (f < Norb: up word of orbital f+1, else dw word of orbital f+1-Norb) with
 * n electrons, as build_Hv_sector computes it on the host: Hs(f)%map and the CSR of spH0ups / spH0dws(iorb) in insertion
 * order.  Returns the dimension; arrays may be NULL. */
int64_t edgpu_selftest_orbs_factor(const edgpu_params *p, int f, int n, int32_t *map, int64_t *nnz, int64_t *rowptr,
                                   int64_t *cols, double *vals);
/* HOST ONLY: the halo tables of ONE rank (lcol2 per list entry; the (peer, slot, own column) triples it stores, sorted by
 * (window, source column); pwin[K+1]).  info[4] = {slots, largest slot count of any rank, triples, list entries}.
 * Returns 0, 1 = capacity too small, -1 = the structured row kernel does not apply. */
int edgpu_selftest_halo_rank(const edgpu_params *p, int ndw, int nranks, int rank, int64_t lr, int64_t tbits_opt, int K,
                             int32_t *info, int32_t *lcol2, int cap_e, int32_t *pdst, int32_t *pslot, int32_t *psrc,
                             int cap_p, int32_t *pwin);
/* GPU: the sharded fast H*v path with `nranks` ranks emulated by `nranks` contexts on ONE device (every rank's halo
 * slab is a plain allocation, the pushes of all ranks run first, no arrival flags); x, y are full host vectors.  Test
 * instrumentation only. */
int edgpu_selftest_sharded_hxv(const edgpu_params *p, int nup, int ndw, int nranks, int64_t srow_lr, int64_t srow_t,
                               int64_t col_cluster, const double *x, double *y);
/* GPU: max |a - b| and max |a| of two device vectors of n doubles (full-size comparisons on the device) */
int edgpu_selftest_dev_maxabsdiff(const double *d_a, const double *d_b, int64_t n, double *maxdiff, double *maxabs);
#ifdef __cplusplus
}
#endif
#endif
