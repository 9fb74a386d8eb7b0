// observables.cu -- local observables and energies of the state kept on the device for the Green's function chains.
//
// Replaces lanc_observables (ED_OBSERVABLES.f90:95-363) and lanc_local_energy (:372-600) for one state at T = 0
// (bath_type normal, ed_total_ud = T, DimPh = 1).  The reference gathers the state on the master rank twice
// (es_return_cvector) and walks it there with bdecomp / c / cdg / binary_search per element; here the state stays
// where it is (every rank keeps its shard) and ONE pass over it produces
//   W(a_up, a_dw) = sum over the states whose impurity occupations are (a_up, a_dw) of |gs|^2
// (4^Norb numbers, deterministic per-block partial tables + one fixed-order final sum, all-reduced over the ranks),
// from which every diagonal quantity of the two routines follows on the host: dens, dens_up/dw, docc, magz, sz2,
// n2, s2tot, Prob, the diagonal of imp_density_matrix, the density-density parts of Epot, Ehartree, Dust, Dund.
// The off-diagonal correlators (inter-orbital density matrix, spin-exchange, pair-hopping; Norb > 1 only) need the
// partner element and are a second, gather-type pass (single rank: configuration C4 is a one-GPU workload).
#include <math.h>
#include <string.h>

#include <vector>

#include "engine.h"

#define OBS_THREADS 256
#define OBS_MAXA 32                      // 2^Norb, Norb <= 5

// W partial tables: block b accumulates tab[b][a_dw][a_up] over its columns (no atomics: fixed order)
template <int NA>
__global__ void __launch_bounds__(OBS_THREADS) k_obs_weights(const double *__restrict__ gs, const int32_t *__restrict__ map_up,
                                                             const int32_t *__restrict__ map_dw, int64_t dimup, int64_t qdw,
                                                             int64_t coloff, double *__restrict__ tab) {
  __shared__ double sh[OBS_THREADS / 32][NA];
  __shared__ double mine[NA * NA];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < NA * NA; i += blockDim.x) mine[i] = 0.0;
  __syncthreads();
  for (int64_t jl = blockIdx.x; jl < qdw; jl += gridDim.x) {
    const int ad = map_dw[coloff + jl] & (NA - 1);
    double acc[NA];
#pragma unroll
    for (int a = 0; a < NA; a++) acc[a] = 0.0;
    for (int64_t i = threadIdx.x; i < dimup; i += blockDim.x) {
      const double v = gs[i + jl * dimup];
      const double w = fabs(v) * fabs(v);
      const int au = map_up[i] & (NA - 1);
#pragma unroll
      for (int a = 0; a < NA; a++) acc[a] += (au == a) ? w : 0.0;
    }
#pragma unroll
    for (int a = 0; a < NA; a++) {
      double s = acc[a];
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) sh[warp][a] = s;
    }
    __syncthreads();
    if (threadIdx.x < NA) {
      double s = 0.0;
      for (int w = 0; w < OBS_THREADS / 32; w++) s += sh[w][threadIdx.x];
      mine[ad * NA + threadIdx.x] += s;
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < NA * NA; i += blockDim.x) tab[(size_t)blockIdx.x * NA * NA + i] = mine[i];
}
__global__ void k_obs_sum_tables(const double *__restrict__ tab, int nblocks, int n, double *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; b++) s += tab[(size_t)b * n + i];
  out[i] = s;
}

// off-diagonal correlators (single rank): out[0..24] up hops <c+_io c_jo>, [25..49] dw hops, [50] spin-exchange, [51] pair-hopping
#define OBS_NOFF 52
__global__ void __launch_bounds__(OBS_THREADS) k_obs_offdiag(const double *__restrict__ gs, const int32_t *__restrict__ map_up,
                                                             const int32_t *__restrict__ map_dw, int64_t dimup, int64_t dimdw, int norb,
                                                             const uint32_t *__restrict__ binom, double *__restrict__ tab) {
  __shared__ double sh[OBS_THREADS / 32];
  double acc[OBS_NOFF];
  for (int q = 0; q < OBS_NOFF; q++) acc[q] = 0.0;
  const int64_t dim = dimup * dimdw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < dim; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t iup = i % dimup, idw = i / dimup;
    const uint32_t mup = (uint32_t)map_up[iup], mdw = (uint32_t)map_dw[idw];
    const double gi = gs[i];
    for (int io = 0; io < norb; io++)
      for (int jo = 0; jo < norb; jo++) {
        if (io == jo) continue;
        const uint32_t bi = 1u << io, bj = 1u << jo;
        const bool up_ok = (mup & bj) && !(mup & bi), dw_ok = (mdw & bj) && !(mdw & bi);
        // c(jo) then cdg(io) on one spin word: sign = sign_below(m, jo) * sign_below(m without jo, io)
        if (up_ok) {
          const uint32_t k1 = mup & ~bj, k2 = k1 | bi;
          const double sg = hd_sign_below(mup, jo + 1) * hd_sign_below(k1, io + 1);
          acc[io + 5 * jo] += sg * gi * gs[hd_rank(k2, binom) + idw * dimup];
        }
        if (dw_ok) {
          const uint32_t k1 = mdw & ~bj, k2 = k1 | bi;
          const double sg = hd_sign_below(mdw, jo + 1) * hd_sign_below(k1, io + 1);
          acc[25 + io + 5 * jo] += sg * gi * gs[iup + hd_rank(k2, binom) * dimup];
        }
        // spin-exchange (:449-470): dw io -> jo, up jo -> io
        if ((mup & bj) && (mdw & bi) && !(mdw & bj) && !(mup & bi)) {
          const uint32_t d1 = mdw & ~bi, d2 = d1 | bj, u1 = mup & ~bj, u2 = u1 | bi;
          const double sg = hd_sign_below(mdw, io + 1) * hd_sign_below(d1, jo + 1) * hd_sign_below(mup, jo + 1) * hd_sign_below(u1, io + 1);
          acc[50] += sg * gi * gs[hd_rank(u2, binom) + hd_rank(d2, binom) * dimup];
        }
        // pair-hopping (:475-496): dw jo -> io, up jo -> io
        if (up_ok && dw_ok) {
          const uint32_t d1 = mdw & ~bj, d2 = d1 | bi, u1 = mup & ~bj, u2 = u1 | bi;
          const double sg = hd_sign_below(mdw, jo + 1) * hd_sign_below(d1, io + 1) * hd_sign_below(mup, jo + 1) * hd_sign_below(u1, io + 1);
          acc[51] += sg * gi * gs[hd_rank(u2, binom) + hd_rank(d2, binom) * dimup];
        }
      }
  }
  for (int q = 0; q < OBS_NOFF; q++) {
    double s = acc[q];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < OBS_THREADS / 32; w++) t += sh[w];
      tab[(size_t)blockIdx.x * OBS_NOFF + q] = t;
    }
    __syncthreads();
  }
}

template <int NA>
static void launch_weights(int grid, cudaStream_t st, const double *gs, const int32_t *mu, const int32_t *md, int64_t dimup, int64_t qdw,
                           int64_t coloff, double *tab) {
  k_obs_weights<NA><<<grid, OBS_THREADS, 0, st>>>(gs, mu, md, dimup, qdw, coloff, tab);
}

extern "C" int edgpu_observables_normal(edgpu_ctx *c, double zeta, edgpu_observables *out) {
  if (!c || !out) return edgpu_set_err(EDGPU_ERR_INVALID, "observables: bad arguments");
  if (!c->hp.ed_total_ud) return edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "ed_total_ud = F: observables of an orbital-resolved state are not built");
  if (!c->d_gs) return edgpu_set_err(EDGPU_ERR_INVALID, "observables: no state set (edgpu_gf_set_state / edgpu_gf_set_state_from_eigh)");
  if (!(zeta > 0.0)) return edgpu_set_err(EDGPU_ERR_INVALID, "observables: zeta_function must be positive");
  CK(cudaSetDevice(c->device));
  const int norb = c->dp.norb, NA = 1 << norb, LD = EDGPU_MAX_ORB;
  const int64_t dimup = c->h_binom[c->ns * EDGPU_BINOM_LD + c->gs_nup], dimdw = c->h_binom[c->ns * EDGPU_BINOM_LD + c->gs_ndw];
  int64_t qdw, coloff;
  edgpu_split(dimdw, c->nranks, c->rank, &qdw, &coloff);
  if (dimup * qdw != c->gs_nloc) return edgpu_set_err(EDGPU_ERR_INVALID, "observables: state shard does not match the rank layout");
  int32_t *d_mu = nullptr, *d_md = nullptr;
  double *d_tab = nullptr, *d_w = nullptr;
  int rc = build_sector_map_device(c, c->gs_nup, &d_mu);
  if (!rc) rc = build_sector_map_device(c, c->gs_ndw, &d_md);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(qdw, (int64_t)c->sm_count * 4));
  std::vector<double> W((size_t)NA * NA, 0.0), off(OBS_NOFF, 0.0);
  if (!rc && cudaMalloc(&d_tab, (size_t)std::max(grid * NA * NA, c->sm_count * 4 * OBS_NOFF) * sizeof(double)) != cudaSuccess) rc = edgpu_set_err(EDGPU_ERR_CUDA, "observables: cudaMalloc");
  if (!rc && cudaMalloc(&d_w, (size_t)std::max(NA * NA, OBS_NOFF) * sizeof(double)) != cudaSuccess) rc = edgpu_set_err(EDGPU_ERR_CUDA, "observables: cudaMalloc");
  if (!rc) {
    switch (norb) {
      case 1: launch_weights<2>(grid, c->stream, c->d_gs, d_mu, d_md, dimup, qdw, coloff, d_tab); break;
      case 2: launch_weights<4>(grid, c->stream, c->d_gs, d_mu, d_md, dimup, qdw, coloff, d_tab); break;
      case 3: launch_weights<8>(grid, c->stream, c->d_gs, d_mu, d_md, dimup, qdw, coloff, d_tab); break;
      case 4: launch_weights<16>(grid, c->stream, c->d_gs, d_mu, d_md, dimup, qdw, coloff, d_tab); break;
      default: launch_weights<32>(grid, c->stream, c->d_gs, d_mu, d_md, dimup, qdw, coloff, d_tab); break;
    }
    c->launches++;
    k_obs_sum_tables<<<(NA * NA + 255) / 256, 256, 0, c->stream>>>(d_tab, grid, NA * NA, d_w);
    c->launches++;
    if (cudaGetLastError() != cudaSuccess) rc = edgpu_set_err(EDGPU_ERR_CUDA, "observables: kernel launch");
  }
  if (!rc) rc = comm_allreduce_array(c, d_w, NA * NA);
  if (!rc && cudaMemcpyAsync(W.data(), d_w, W.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) rc = edgpu_set_err(EDGPU_ERR_CUDA, "observables: read-back");
  if (!rc && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = edgpu_set_err(EDGPU_ERR_CUDA, "observables: sync");
  if (!rc && norb > 1) {
    if (c->nranks > 1) rc = edgpu_set_err(EDGPU_ERR_UNSUPPORTED, "observables: inter-orbital correlators of a sharded state are not implemented (multi-orbital models are one-GPU workloads)");
    if (!rc) {
      const int g2 = c->sm_count * 4;
      k_obs_offdiag<<<g2, OBS_THREADS, 0, c->stream>>>(c->d_gs, d_mu, d_md, dimup, dimdw, norb, c->d_binom, d_tab);
      c->launches++;
      k_obs_sum_tables<<<1, 64, 0, c->stream>>>(d_tab, g2, OBS_NOFF, d_w);
      c->launches++;
      if (cudaMemcpyAsync(off.data(), d_w, OBS_NOFF * sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
          cudaStreamSynchronize(c->stream) != cudaSuccess) rc = edgpu_set_err(EDGPU_ERR_CUDA, "observables: off-diagonal pass");
    }
  }
  cudaFree(d_mu); cudaFree(d_md); cudaFree(d_tab); cudaFree(d_w);
  if (rc) return rc;

  // ---- host arithmetic on the 4^Norb weights (the reference's per-state sums, regrouped) ----
  memset(out, 0, sizeof(*out));
  const double peso = 1.0 / zeta;
  const DevParams &P = c->dp;
  for (int ad = 0; ad < NA; ad++)
    for (int au = 0; au < NA; au++) {
      const double w = peso * W[(size_t)ad * NA + au];
      if (w == 0.0) continue;
      double nup[EDGPU_MAX_ORB], ndw[EDGPU_MAX_ORB], sz[EDGPU_MAX_ORB], nt[EDGPU_MAX_ORB], ssum = 0.0;
      int iprob = 0, p3 = 1;
      for (int io = 0; io < norb; io++) {
        nup[io] = (au >> io) & 1; ndw[io] = (ad >> io) & 1;
        sz[io] = (nup[io] - ndw[io]) / 2.0; nt[io] = nup[io] + ndw[io];
        iprob += (int)nt[io] * p3; p3 *= 3;
        ssum += sz[io];
      }
      out->prob[iprob] += w;
      out->s2tot += ssum * ssum * w;
      for (int io = 0; io < norb; io++) {
        out->dens[io] += nt[io] * w; out->dens_up[io] += nup[io] * w; out->dens_dw[io] += ndw[io] * w;
        out->docc[io] += nup[io] * ndw[io] * w; out->magz[io] += (nup[io] - ndw[io]) * w;
        for (int jo = 0; jo < norb; jo++) { out->sz2[io + LD * jo] += sz[io] * sz[jo] * w; out->n2[io + LD * jo] += nt[io] * nt[jo] * w; }
        out->dm[0][io + LD * io] += nup[io] * w;
        if (P.nspin == 2) out->dm[1][io + LD * io] += ndw[io] * w;
        out->eknot += (P.hloc_up[io * EDGPU_MAX_ORB + io] * nup[io] + P.hloc_dw[io * EDGPU_MAX_ORB + io] * ndw[io]) * w;
        out->epot += P.uloc[io] * nup[io] * ndw[io] * w;
        if (P.hfmode) out->ehartree += (-0.5 * P.uloc[io] * (nup[io] + ndw[io]) + 0.25 * P.uloc[io]) * w;
        for (int jo = io + 1; jo < norb; jo++) {
          const double a = nup[io] * ndw[jo] + nup[jo] * ndw[io], b = nup[io] * nup[jo] + ndw[io] * ndw[jo];
          out->epot += (P.ust * a + (P.ust - P.jh) * b) * w;
          out->dust += a * w; out->dund += b * w;
          if (P.hfmode) {
            const double n4 = nup[io] + ndw[io] + nup[jo] + ndw[jo];
            out->ehartree += (-0.5 * P.ust * n4 + 0.25 * P.ust - 0.5 * (P.ust - P.jh) * n4 + 0.25 * (P.ust - P.jh)) * w;
          }
        }
      }
    }
  if (norb > 1) {
    for (int io = 0; io < norb; io++)
      for (int jo = 0; jo < norb; jo++) {
        if (io == jo) continue;
        out->dm[0][io + LD * jo] = peso * off[(size_t)(io + 5 * jo)];
        if (P.nspin == 2) out->dm[1][io + LD * jo] = peso * off[(size_t)(25 + io + 5 * jo)];
        // impHloc(1,1,iorb,jorb) /= 0 terms of <H_imp> (:430-446); hloc_dw is the Nspin-th spin block
        out->eknot += peso * (P.hloc_up[io * EDGPU_MAX_ORB + jo] * off[(size_t)(io + 5 * jo)] + P.hloc_dw[io * EDGPU_MAX_ORB + jo] * off[(size_t)(25 + io + 5 * jo)]);
      }
    if (P.jhflag && P.jx != 0.0) { out->dse = peso * off[50]; out->epot += P.jx * out->dse; }
    if (P.jhflag && P.jp != 0.0) { out->dph = peso * off[51]; out->epot += P.jp * out->dph; }
  }
  out->epot += out->ehartree;                                      // :587
  return EDGPU_OK;
}
