// hd_funcs.h -- __host__ __device__ integer / bit logic shared by the CUDA kernels and by the
// host-side self-test hooks.  Everything here is a B200-side re-derivation of what the reference
// computes with bdecomp / c / cdg / binary_search loops; file:line citations point at the
// reference routine whose RESULT each function reproduces.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#else
#define HD inline
#endif

#define EDGPU_MAX_ORB 5
#define EDGPU_MAX_SITES 32          // Ns <= 31 so that a spin word fits int32 like the reference
#define EDGPU_MAX_ROW_NNZ 96        // >= Norb^2 + 2*Norb*Nbath (+ Norb^2*Nbath, replica) for every supported model
#define EDGPU_MAX_REPL_BATH 10      // replica baths: Nbath <= 10 and Norb <= 3 (the Hbath matrices travel as kernel parameters)
#define EDGPU_BINOM_LD 33

// Parameters the kernels need, by value (fits the 4 KB kernel-argument space).
struct DevParams {
  int norb, nbath, ns, hfmode, nspin, jhflag;
  int bath_type, nfoo;                             // 0 normal, 1 hybrid, 2 replica; nfoo = size(bath_diag,2): 1 for hybrid
  double uloc[EDGPU_MAX_ORB];
  double ust, jh, jx, jp, xmu;
  double hloc_up[EDGPU_MAX_ORB * EDGPU_MAX_ORB];   // impHloc(1,1,io,jo)       [io*5+jo]
  double hloc_dw[EDGPU_MAX_ORB * EDGPU_MAX_ORB];   // impHloc(Nspin,Nspin,io,jo)
  double be_up[EDGPU_MAX_SITES], be_dw[EDGPU_MAX_SITES];   // bath_diag(1|Nspin, io, kp) [io*nbath+kp]
  double bv_up[EDGPU_MAX_SITES], bv_dw[EDGPU_MAX_SITES];   // diag_hybr(1|Nspin, io, kp)
  // replica: Hbath(1,1,io,jo,kp) / Hbath(Nspin,Nspin,io,jo,kp), [kp*9 + io*3 + jo] (Norb <= 3, Nbath <= 10)
  double hb_up[EDGPU_MAX_REPL_BATH * 9], hb_dw[EDGPU_MAX_REPL_BATH * 9];
};
// 1-based site of bath level kp (0-based) of orbital io (0-based): getBathStride, ED_SETUP.f90:358-375
#define HD_BATH_SITE(P, io, kp) ((P).bath_type == 1 ? (P).norb + (kp) + 1 : ((P).bath_type == 2 ? (io) + 1 + ((kp) + 1) * (P).norb : (P).norb + (io) * (P).nbath + (kp) + 1))

HD int hd_popc(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}

// parity sign of the occupied sites below 1-based site `pos`: the fsgn of c / cdg
// (ED_SETUP.f90:805-831)
HD double hd_sign_below(uint32_t m, int pos) {
  return (hd_popc(m & ((1u << (pos - 1)) - 1u)) & 1) ? -1.0 : 1.0;
}

// 0-based position of state m in the ascending list of all words with popc(m) set bits
// (= binary_search(Hs%map, m) - 1, ED_SETUP.f90:1042-1059; closed form: SURVEY Appendix C).
// binom: row-major table binom[p*EDGPU_BINOM_LD + j] = C(p, j).
HD int64_t hd_rank(uint32_t m, const uint32_t *binom) {
  int64_t r = 0;
  int j = 0;
  while (m) {
#if defined(__CUDA_ARCH__)
    int p = __ffs((int)m) - 1;
#else
    int p = __builtin_ctz(m);
#endif
    j++;
    r += binom[p * EDGPU_BINOM_LD + j];
    m &= m - 1;
  }
  return r;
}

// binary_search with the reference's semantics on a device/host array: 1-based position, 0 if
// absent (ED_SETUP.f90:1042-1059).
HD int64_t hd_binary_search(const int32_t *a, int64_t n, int32_t value) {
  int64_t base = 0;
  while (n > 0) {
    int64_t mid = n / 2 + 1;
    int32_t am = a[base + mid - 1];
    if (am > value) n = mid - 1;
    else if (am < value) { base += mid; n -= mid; }
    else return base + mid;
  }
  return 0;
}

// No-FMA arithmetic so that sums are evaluated exactly in the reference's written order.
HD double hd_mul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  volatile double r = a * b; return r;
#endif
}
HD double hd_add(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  volatile double r = a + b; return r;
#endif
}

#define HD_BIT(m, site1) ((int)(((m) >> ((site1) - 1)) & 1u))

// Diagonal element in the exact summation order of stored/H_local.f90:13-71
// (== direct/HxV_local.f90:15-73).
HD double hd_diag_element(const DevParams &P, uint32_t mup, uint32_t mdw) {
  double h = 0.0;
  for (int io = 0; io < P.norb; io++) {
    int nu = HD_BIT(mup, io + 1), nd = HD_BIT(mdw, io + 1);
    h = hd_add(h, hd_mul(P.hloc_up[io * EDGPU_MAX_ORB + io], (double)nu));
    h = hd_add(h, hd_mul(P.hloc_dw[io * EDGPU_MAX_ORB + io], (double)nd));
    h = hd_add(h, -hd_mul(P.xmu, (double)(nu + nd)));
  }
  for (int io = 0; io < P.norb; io++) {
    int nu = HD_BIT(mup, io + 1), nd = HD_BIT(mdw, io + 1);
    h = hd_add(h, hd_mul(hd_mul(P.uloc[io], (double)nu), (double)nd));
  }
  if (P.norb > 1) {
    for (int io = 0; io < P.norb; io++)
      for (int jo = io + 1; jo < P.norb; jo++) {
        int t = HD_BIT(mup, io + 1) * HD_BIT(mdw, jo + 1) + HD_BIT(mup, jo + 1) * HD_BIT(mdw, io + 1);
        h = hd_add(h, hd_mul(P.ust, (double)t));
      }
    double ujh = hd_add(P.ust, -P.jh);
    for (int io = 0; io < P.norb; io++)
      for (int jo = io + 1; jo < P.norb; jo++) {
        int t = HD_BIT(mup, io + 1) * HD_BIT(mup, jo + 1) + HD_BIT(mdw, io + 1) * HD_BIT(mdw, jo + 1);
        h = hd_add(h, hd_mul(ujh, (double)t));
      }
  }
  if (P.hfmode) {
    for (int io = 0; io < P.norb; io++) {
      int n = HD_BIT(mup, io + 1) + HD_BIT(mdw, io + 1);
      h = hd_add(h, -hd_mul(hd_mul(0.5, P.uloc[io]), (double)n));
      h = hd_add(h, hd_mul(0.25, P.uloc[io]));
    }
    if (P.norb > 1) {
      double ujh = hd_add(P.ust, -P.jh);
      for (int io = 0; io < P.norb; io++)
        for (int jo = io + 1; jo < P.norb; jo++) {
          int n = HD_BIT(mup, io + 1) + HD_BIT(mdw, io + 1) + HD_BIT(mup, jo + 1) + HD_BIT(mdw, jo + 1);
          h = hd_add(h, -hd_mul(hd_mul(0.5, P.ust), (double)n));
          h = hd_add(h, hd_mul(0.25, P.ust));
          h = hd_add(h, -hd_mul(hd_mul(0.5, ujh), (double)n));
          h = hd_add(h, hd_mul(0.25, ujh));
        }
    }
  }
  for (int io = 0; io < P.nfoo; io++)                      // size(bath_diag,2)
    for (int kp = 0; kp < P.nbath; kp++) {
      int site = HD_BATH_SITE(P, io, kp);
      h = hd_add(h, hd_mul(P.be_up[io * P.nbath + kp], (double)HD_BIT(mup, site)));
      h = hd_add(h, hd_mul(P.be_dw[io * P.nbath + kp], (double)HD_BIT(mdw, site)));
    }
  return h;
}

// One hop  c^+_{to} c_{from}  out of source word m (1-based sites): returns the target word and
// the value amp*sg1*sg2 exactly as the reference forms it (c on m, then cdg on the intermediate).
HD uint32_t hd_hop(uint32_t m, int from, int to, double amp, double *val) {
  double sg1 = hd_sign_below(m, from);
  uint32_t k1 = m & ~(1u << (from - 1));
  double sg2 = hd_sign_below(k1, to);
  *val = hd_mul(hd_mul(amp, sg1), sg2);
  return k1 | (1u << (to - 1));
}

// Row `t` (target word) of the one-spin factor spH0ups(1)/spH0dws(1): all (source word, value)
// pairs that stored/H_up.f90:8-81 / H_dw.f90:8-80 would insert at (row=target, col=source), in
// the reference's per-source enumeration order (impHloc loop, then hybridisation loop).
// spin: 0 = up (impHloc(1,1), diag_hybr(1)), 1 = dw (index Nspin).  Returns the entry count.
HD int hd_factor_row_sources(const DevParams &P, int spin, uint32_t t, uint32_t *src, double *val) {
  const double *hloc = spin ? P.hloc_dw : P.hloc_up;
  const double *bv = spin ? P.bv_dw : P.bv_up;
  int n = 0;
  for (int io = 0; io < P.norb; io++)
    for (int jo = 0; jo < P.norb; jo++) {
      double a = hloc[io * EDGPU_MAX_ORB + jo];
      // forward: source has n[jo]=1, n[io]=0  ->  target has n[jo]=0, n[io]=1
      if (io != jo && a != 0.0 && HD_BIT(t, io + 1) == 1 && HD_BIT(t, jo + 1) == 0) {
        uint32_t m = (t & ~(1u << io)) | (1u << jo);
        double v;
        hd_hop(m, jo + 1, io + 1, a, &v);
        src[n] = m; val[n] = v; n++;
      }
    }
  if (P.bath_type == 2) {                                    // replica inter-orbital bath hopping, stored/H_up.f90:26-50
    const double *hb = spin ? P.hb_dw : P.hb_up;
    for (int kp = 0; kp < P.nbath; kp++)
      for (int io = 0; io < P.norb; io++)
        for (int jo = 0; jo < P.norb; jo++) {
          double a = hb[kp * 9 + io * 3 + jo];
          int ialfa = HD_BATH_SITE(P, io, kp), ibeta = HD_BATH_SITE(P, jo, kp);
          // source n[ibeta]=1, n[ialfa]=0  ->  target n[ibeta]=0, n[ialfa]=1
          if (io != jo && a != 0.0 && HD_BIT(t, ialfa) == 1 && HD_BIT(t, ibeta) == 0) {
            uint32_t m = (t & ~(1u << (ialfa - 1))) | (1u << (ibeta - 1));
            double v;
            hd_hop(m, ibeta, ialfa, a, &v);
            src[n] = m; val[n] = v; n++;
          }
        }
  }
  for (int io = 0; io < P.norb; io++)
    for (int kp = 0; kp < P.nbath; kp++) {
      double a = bv[io * P.nbath + kp];
      if (a == 0.0) continue;
      int ialfa = HD_BATH_SITE(P, io, kp);
      // source n[io]=1,n[ialfa]=0 (c(io), cdg(ialfa))  ->  target n[io]=0,n[ialfa]=1
      if (HD_BIT(t, io + 1) == 0 && HD_BIT(t, ialfa) == 1) {
        uint32_t m = (t | (1u << io)) & ~(1u << (ialfa - 1));
        double v;
        hd_hop(m, io + 1, ialfa, a, &v);
        src[n] = m; val[n] = v; n++;
      }
      // source n[io]=0,n[ialfa]=1 (c(ialfa), cdg(io))  ->  target n[io]=1,n[ialfa]=0
      if (HD_BIT(t, io + 1) == 1 && HD_BIT(t, ialfa) == 0) {
        uint32_t m = (t & ~(1u << io)) | (1u << (ialfa - 1));
        double v;
        hd_hop(m, ialfa, io + 1, a, &v);
        src[n] = m; val[n] = v; n++;
      }
    }
  return n;
}

// One row of spH0ups/spH0dws: sources -> column positions by device binary search, ordered by
// ascending source (the reference's outer loop is over sources), duplicates accumulated.
HD int hd_factor_row(const DevParams &P, int spin, const int32_t *map, int64_t n, uint32_t t,
                          int32_t *cols, double *vals) {
  uint32_t src[EDGPU_MAX_ROW_NNZ];
  int cnt = hd_factor_row_sources(P, spin, t, src, vals);
  for (int k = 0; k < cnt; k++) cols[k] = (int32_t)(hd_binary_search(map, n, (int32_t)src[k]) - 1);
  for (int a = 1; a < cnt; a++) {                           // stable insertion sort by column
    int32_t cc = cols[a]; double vv = vals[a];
    int b = a - 1;
    while (b >= 0 && cols[b] > cc) { cols[b + 1] = cols[b]; vals[b + 1] = vals[b]; b--; }
    cols[b + 1] = cc; vals[b + 1] = vv;
  }
  int m = 0;
  for (int a = 0; a < cnt; a++) {                           // sp_insert_element: accumulate repeats
    if (m > 0 && cols[m - 1] == cols[a]) vals[m - 1] = hd_add(vals[m - 1], vals[a]);
    else { cols[m] = cols[a]; vals[m] = vals[a]; m++; }
  }
  return m;
}
// Row (mup,mdw) of spH0nd in insertion order (stored/H_non_local.f90:21-83): spin-exchange
// entries over (iorb,jorb), then pair-hopping.  Outputs the column's (up word, dw word).
HD int hd_nonlocal_row(const DevParams &P, uint32_t mup, uint32_t mdw, uint32_t *cup, uint32_t *cdw,
                       double *val) {
  int n = 0;
  if (!P.jhflag) return 0;
  if (P.jx != 0.0)
    for (int io = 0; io < P.norb; io++)
      for (int jo = 0; jo < P.norb; jo++)
        if (io != jo && HD_BIT(mup, jo + 1) == 1 && HD_BIT(mdw, io + 1) == 1 &&
            HD_BIT(mdw, jo + 1) == 0 && HD_BIT(mup, io + 1) == 0) {
          double sg1 = hd_sign_below(mdw, io + 1);
          uint32_t k1 = mdw & ~(1u << io);
          double sg2 = hd_sign_below(k1, jo + 1);
          uint32_t k2 = k1 | (1u << jo);
          double sg3 = hd_sign_below(mup, jo + 1);
          uint32_t k3 = mup & ~(1u << jo);
          double sg4 = hd_sign_below(k3, io + 1);
          uint32_t k4 = k3 | (1u << io);
          cup[n] = k4; cdw[n] = k2;
          val[n] = hd_mul(hd_mul(hd_mul(hd_mul(P.jx, sg1), sg2), sg3), sg4);
          n++;
        }
  if (P.jp != 0.0)
    for (int io = 0; io < P.norb; io++)
      for (int jo = 0; jo < P.norb; jo++)
        if (HD_BIT(mup, jo + 1) == 1 && HD_BIT(mdw, jo + 1) == 1 && HD_BIT(mdw, io + 1) == 0 &&
            HD_BIT(mup, io + 1) == 0) {
          double sg1 = hd_sign_below(mdw, jo + 1);
          uint32_t k1 = mdw & ~(1u << jo);
          double sg2 = hd_sign_below(k1, io + 1);
          uint32_t k2 = k1 | (1u << io);
          double sg3 = hd_sign_below(mup, jo + 1);
          uint32_t k3 = mup & ~(1u << jo);
          double sg4 = hd_sign_below(k3, io + 1);
          uint32_t k4 = k3 | (1u << io);
          cup[n] = k4; cdw[n] = k2;
          val[n] = hd_mul(hd_mul(hd_mul(hd_mul(P.jp, sg1), sg2), sg3), sg4);
          n++;
        }
  return n;
}

// Factorised diagonal for the on-the-fly ("direct") operator:
//   Hd(iup,idw) = dfac_up[iup] + dfac_dw[idw] + sum_{a,b} W_ab n_up,a n_dw,b
// with W_aa = Uloc(a), W_ab = Ust (a/=b).  Same terms as stored/H_local.f90, regrouped so that
// only O(DimUp + DimDw) values are tabulated (16 B/element traffic instead of 24).
HD double hd_diag_factor(const DevParams &P, int spin, uint32_t m) {
  const double *hloc = spin ? P.hloc_dw : P.hloc_up;
  const double *be = spin ? P.be_dw : P.be_up;
  double h = 0.0;
  double ujh = P.ust - P.jh;
  for (int io = 0; io < P.norb; io++) {
    int n = HD_BIT(m, io + 1);
    h += (hloc[io * EDGPU_MAX_ORB + io] - P.xmu) * n;
    if (P.hfmode) h += -0.5 * P.uloc[io] * n + (spin ? 0.0 : 0.25 * P.uloc[io]);
    for (int jo = io + 1; jo < P.norb; jo++) {
      int nj = HD_BIT(m, jo + 1);
      h += ujh * (n * nj);
      if (P.hfmode) {
        h += -0.5 * P.ust * (n + nj) - 0.5 * ujh * (n + nj);
        if (!spin) h += 0.25 * P.ust + 0.25 * ujh;
      }
    }
  }
  for (int io = 0; io < P.nfoo; io++)
    for (int kp = 0; kp < P.nbath; kp++)
      h += be[io * P.nbath + kp] * HD_BIT(m, HD_BATH_SITE(P, io, kp));
  return h;
}
HD double hd_diag_cross(const DevParams &P, uint32_t mup, uint32_t mdw) {
  double h = 0.0;
  for (int a = 0; a < P.norb; a++) {
    if (!HD_BIT(mup, a + 1)) continue;
    for (int b = 0; b < P.norb; b++)
      if (HD_BIT(mdw, b + 1)) h += (a == b) ? P.uloc[a] : P.ust;
  }
  return h;
}
