/*
 * ed_oracle.c -- CPU ORACLE (test infrastructure only; see ed_oracle.h).
 *
 * PARITY UNPINNED (no reference golden vectors exist, reference not compilable here).
 *
 * Loop-order-faithful C restatement of the N_up:N_dw Lanczos H*v path of
 * lcrippa/dmft-lanc-ed.  Compile with -ffp-contract=off so that floating-point sums are
 * evaluated exactly in the written order (the reference is built with gfortran -O3
 * -funroll-loops and no -march, CMakeLists.txt:135, i.e. without FMA contraction on x86-64).
 *
 * The Lanczos recurrence itself lives in SciFortran (SF_SP_LINALG: sp_lanc_eigh /
 * sp_lanc_tridiag), an un-vendored, un-pinned dependency (CMakeLists.txt:91-106).  It is
 * restated here from its published algorithm (plain three-term Lanczos, tridiagonal QL);
 * parity for it is anchored on the reference's call sites ED_DIAG.f90:174-186 and
 * ED_GF_NORMAL.f90:232-237.
 */
#include "ed_oracle.h"
#include <complex.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define HLOC(c, is, js, io, jo) \
  ((c)->imphloc[(is) + (c)->nspin * ((js) + (c)->nspin * ((io) + (c)->norb * (jo)))])
#define BATHE(c, is, io, ib) ((c)->bath_e[(is) + (c)->nspin * ((io) + (c)->norb * (ib))])
#define BATHV(c, is, io, ib) ((c)->bath_v[(is) + (c)->nspin * ((io) + (c)->norb * (ib))])

static void *xmalloc(size_t n) {
  void *p = malloc(n ? n : 1);
  if (!p) { fprintf(stderr, "ed_oracle: out of memory (%zu bytes)\n", n); abort(); }
  return p;
}
static void *xcalloc(size_t n, size_t s) {
  void *p = calloc(n ? n : 1, s ? s : 1);
  if (!p) { fprintf(stderr, "ed_oracle: out of memory\n"); abort(); }
  return p;
}


/* Minimal parallel-for over emulated MPI ranks (pthreads; the image has no libgomp).  Rank r
 * is executed by thread r % nthreads, i.e. one thread per rank when nthreads >= nranks. */
typedef void (*rank_fn)(int r, void *arg);
typedef struct { rank_fn fn; void *arg; int t, nthreads, nranks; } par_job;
static void *par_worker(void *p) {
  par_job *j = (par_job *)p;
  for (int r = j->t; r < j->nranks; r += j->nthreads) j->fn(r, j->arg);
  return NULL;
}
static void par_for(int nthreads, int nranks, rank_fn fn, void *arg) {
  if (nthreads > nranks) nthreads = nranks;
  if (nthreads <= 1) { for (int r = 0; r < nranks; r++) fn(r, arg); return; }
  pthread_t *th = (pthread_t *)xmalloc((size_t)nthreads * sizeof(pthread_t));
  par_job *jb = (par_job *)xmalloc((size_t)nthreads * sizeof(par_job));
  for (int t = 0; t < nthreads; t++) {
    jb[t].fn = fn; jb[t].arg = arg; jb[t].t = t; jb[t].nthreads = nthreads; jb[t].nranks = nranks;
    pthread_create(&th[t], NULL, par_worker, &jb[t]);
  }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th); free(jb);
}

/* ===================================================================================== */
/* setup                                                                                 */
/* ===================================================================================== */
orc_ctx *orc_ctx_create(int norb, int nbath, int nspin, int hfmode, const double *uloc,
                        double ust, double jh, double jx, double jp, double xmu,
                        const double *imphloc, const double *bath_e, const double *bath_v) {
  orc_ctx *c = (orc_ctx *)xcalloc(1, sizeof(orc_ctx));
  c->norb = norb; c->nbath = nbath; c->nspin = nspin; c->hfmode = hfmode;
  c->ns = (nbath + 1) * norb;                         /* ED_SETUP.f90:113-116 (normal bath) */
  for (int i = 0; i < ORC_MAX_ORB; i++) c->uloc[i] = (i < norb) ? uloc[i] : 0.0;
  c->ust = ust; c->jh = jh; c->jx = jx; c->jp = jp; c->xmu = xmu;
  c->jhflag = (norb > 1 && (jx != 0.0 || jp != 0.0)); /* ED_SETUP.f90:147-148 */
  size_t nh = (size_t)nspin * nspin * norb * norb, nb = (size_t)nspin * norb * nbath;
  c->imphloc = (double *)xcalloc(nh, sizeof(double));
  c->bath_e = (double *)xcalloc(nb, sizeof(double));
  c->bath_v = (double *)xcalloc(nb, sizeof(double));
  if (imphloc) memcpy(c->imphloc, imphloc, nh * sizeof(double));
  memcpy(c->bath_e, bath_e, nb * sizeof(double));
  memcpy(c->bath_v, bath_v, nb * sizeof(double));
  c->bath_type = 0; c->nfoo = norb; c->bath_h = NULL;
  return c;
}
orc_ctx *orc_ctx_create_bt(int norb, int nbath, int nspin, int hfmode, const double *uloc,
                           double ust, double jh, double jx, double jp, double xmu, const double *imphloc,
                           int bath_type, const double *bath_e, const double *bath_v, const double *bath_h) {
  if (bath_type == 0) return orc_ctx_create(norb, nbath, nspin, hfmode, uloc, ust, jh, jx, jp, xmu, imphloc, bath_e, bath_v);
  size_t nb = (size_t)nspin * norb * nbath;
  double *e = (double *)xcalloc(nb, sizeof(double)), *v = (double *)xcalloc(nb, sizeof(double));
  for (int is = 0; is < nspin; is++)
    for (int io = 0; io < norb; io++)
      for (int kp = 0; kp < nbath; kp++) {
        size_t at = (size_t)is + nspin * ((size_t)io + norb * (size_t)kp);
        if (bath_type == 1) {                              /* hybrid */
          v[at] = bath_v[at];
          if (io == 0) e[at] = bath_e[is + nspin * kp];
        } else {                                           /* replica, ED_HAMILTONIAN_SPARSE_HxV.f90:61-75 */
          v[at] = bath_v[is + nspin * kp];
          e[at] = bath_h[is + nspin * (is + nspin * (io + norb * (io + norb * (size_t)kp)))];
        }
      }
  orc_ctx *c = orc_ctx_create(norb, nbath, nspin, hfmode, uloc, ust, jh, jx, jp, xmu, imphloc, e, v);
  free(e); free(v);
  c->bath_type = bath_type;
  if (bath_type == 1) { c->ns = norb + nbath; c->nfoo = 1; }
  if (bath_type == 2) {
    size_t nh = (size_t)nspin * nspin * norb * norb * nbath;
    c->bath_h = (double *)xcalloc(nh, sizeof(double));
    memcpy(c->bath_h, bath_h, nh * sizeof(double));
  }
  return c;
}
void orc_ctx_destroy(orc_ctx *c) {
  if (!c) return;
  free(c->imphloc); free(c->bath_e); free(c->bath_v); free(c->bath_h); free(c);
}

/* init_dmft_bath, bath_type normal: ED_BATH/dmft_aux.f90:102-133.  Arrays (nspin,norb,nbath). */
void orc_init_dmft_bath(int norb, int nbath, int nspin, double hwband, double *e, double *v) {
  int so = nspin * norb;
#define SETE(ib, val) do { for (int q = 0; q < so; q++) e[q + so * (ib)] = (val); } while (0)
  SETE(0, -hwband);
  SETE(nbath - 1, hwband);
  int nh = nbath / 2;
  if (nbath % 2 == 0 && nbath >= 4) {
    double de = hwband / (double)((nh - 1) > 1 ? (nh - 1) : 1);
    SETE(nh - 1, -1.0e-1);
    SETE(nh, 1.0e-1);
    for (int i = 2; i <= nh - 1; i++) {
      SETE(i - 1, -hwband + (i - 1) * de);
      SETE(nbath - i, hwband - (i - 1) * de);
    }
  } else if (nbath % 2 != 0 && nbath >= 3) {
    double de = hwband / (double)nh;
    SETE(nh, 0.0);
    for (int i = 2; i <= nh; i++) {
      SETE(i - 1, -hwband + (i - 1) * de);
      SETE(nbath - i, hwband - (i - 1) * de);
    }
  }
#undef SETE
  double vv = 1.0 / sqrt((double)nbath);
  if (vv < 0.1) vv = 0.1;
  for (int ib = 0; ib < nbath; ib++)
    for (int q = 0; q < so; q++) v[q + so * ib] = vv;
}

/* ===================================================================================== */
/* ED_SETUP hot subset                                                                   */
/* ===================================================================================== */
/* binomial, ED_SETUP.f90:1017-1035 (floating-point product, rounded) */
int orc_binomial(int n1, int n2) {
  if (n2 < 0) return 0;
  if (n2 == 0) return 1;
  double xh = 1.0;
  for (int i = 1; i <= n2; i++) xh = xh * (double)(n1 + 1 - i) / (double)i;
  return (int)(xh + 0.5);
}

/* build_sector (one spin species), ED_SETUP.f90:764-777: ascending scan with popcnt filter */
int64_t orc_build_sector_map(int ns, int n, int32_t *map) {
  int64_t dim = 0;
  int64_t top = (int64_t)1 << ns;
  for (int64_t s = 0; s < top; s++) {
    if (__builtin_popcountll((unsigned long long)s) != n) continue;
    if (map) map[dim] = (int32_t)s;
    dim++;
  }
  return dim;
}

/* c, ED_SETUP.f90:805-817.  pos is the 1-based site.  Returns 0 on success, 1 if empty. */
int orc_c(int pos, int32_t in, int32_t *out, double *fsgn) {
  if (!((in >> (pos - 1)) & 1)) return 1;
  double s = 1.0;
  for (int l = 1; l <= pos - 1; l++)
    if ((in >> (l - 1)) & 1) s = -s;
  *fsgn = s;
  *out = in & ~((int32_t)1 << (pos - 1));
  return 0;
}
/* cdg, ED_SETUP.f90:819-831 */
int orc_cdg(int pos, int32_t in, int32_t *out, double *fsgn) {
  if ((in >> (pos - 1)) & 1) return 1;
  double s = 1.0;
  for (int l = 1; l <= pos - 1; l++)
    if ((in >> (l - 1)) & 1) s = -s;
  *fsgn = s;
  *out = in | ((int32_t)1 << (pos - 1));
  return 0;
}

/* binary_search, ED_SETUP.f90:1042-1059: 1-based position in the ascending list, 0 = absent.
 * Same probe sequence (mid = size/2 + 1) written as a loop instead of recursion on slices. */
int64_t orc_binary_search(const int32_t *a, int64_t n, int32_t value) {
  int64_t base = 0;
  while (n > 0) {
    int64_t mid = n / 2 + 1;                 /* 1-based within the current slice */
    int32_t am = a[base + mid - 1];
    if (am > value) {
      n = mid - 1;
    } else if (am < value) {
      base += mid;
      n -= mid;
    } else {
      return base + mid;
    }
  }
  return 0;
}

/* get_Sector for ed_total_ud=T, QN=[Nup,Ndw]: ED_SETUP.f90:446-457 */
int orc_get_sector(int nup, int ndw, int ns) { return 1 + ndw + nup * (ns + 1); }
/* get_Nup/get_Ndw, ED_SETUP.f90:477-500 */
void orc_get_nup_ndw(int isector, int ns, int *nup, int *ndw) {
  int count = isector - 1;
  *ndw = count % (ns + 1);
  *nup = count / (ns + 1);
}
/* getBathStride, ED_SETUP.f90:358-375.  iorb,kp 1-based; returns 1-based site */
int orc_bath_stride(const orc_ctx *c, int iorb, int kp) {
  if (c->bath_type == 1) return c->norb + kp;                /* hybrid */
  if (c->bath_type == 2) return iorb + kp * c->norb;         /* replica */
  return c->norb + (iorb - 1) * c->nbath + kp;               /* normal */
}

static void bdecomp(int32_t m, int ns, int *ivec) {        /* ED_SETUP.f90:937-947 */
  for (int l = 0; l < ns; l++) ivec[l] = (m >> l) & 1;
}

/* ===================================================================================== */
/* ED_SPARSE_MATRIX: row lists with insert-or-accumulate                                 */
/* ===================================================================================== */
typedef struct { int size, cap; int64_t *cols; double *vals; } rl_row;
typedef struct { int64_t nrow, ncol; rl_row *row; } rl_mat;

static void rl_init(rl_mat *m, int64_t n) {                /* sp_init_matrix, :118-146 */
  m->nrow = n; m->ncol = n;
  m->row = (rl_row *)xcalloc((size_t)n, sizeof(rl_row));
}
/* sp_insert_element, ED_SPARSE_MATRIX.f90:255-285: accumulate if the column is already in
 * the row, else append.  (The reference locates an existing column with binary_search on the
 * row's column list; a linear find returns the same position whenever that search succeeds.) */
static void rl_insert(rl_mat *m, double value, int64_t i, int64_t j) {
  rl_row *r = &m->row[i];
  for (int p = 0; p < r->size; p++)
    if (r->cols[p] == j) { r->vals[p] = r->vals[p] + value; return; }
  if (r->size == r->cap) {
    r->cap = r->cap ? 2 * r->cap : 4;
    r->cols = (int64_t *)realloc(r->cols, (size_t)r->cap * sizeof(int64_t));
    r->vals = (double *)realloc(r->vals, (size_t)r->cap * sizeof(double));
    if (!r->cols || !r->vals) abort();
  }
  r->cols[r->size] = j; r->vals[r->size] = value; r->size++;
  if (r->size > m->ncol) { fprintf(stderr, "sp_insert_element ERROR: row%%Size > Ncol\n"); abort(); }
}
static void rl_to_csr(rl_mat *m, orc_csr *o) {             /* Appendix C: insertion order */
  o->nrow = m->nrow; o->ncol = m->ncol;
  o->rowptr = (int64_t *)xmalloc((size_t)(m->nrow + 1) * sizeof(int64_t));
  o->rowptr[0] = 0;
  for (int64_t i = 0; i < m->nrow; i++) o->rowptr[i + 1] = o->rowptr[i] + m->row[i].size;
  int64_t nnz = o->rowptr[m->nrow];
  o->cols = (int64_t *)xmalloc((size_t)nnz * sizeof(int64_t));
  o->vals = (double *)xmalloc((size_t)nnz * sizeof(double));
  for (int64_t i = 0; i < m->nrow; i++) {
    memcpy(o->cols + o->rowptr[i], m->row[i].cols, (size_t)m->row[i].size * sizeof(int64_t));
    memcpy(o->vals + o->rowptr[i], m->row[i].vals, (size_t)m->row[i].size * sizeof(double));
    free(m->row[i].cols); free(m->row[i].vals);
  }
  free(m->row); m->row = NULL;
}
static void csr_free(orc_csr *o) { free(o->rowptr); free(o->cols); free(o->vals); memset(o, 0, sizeof(*o)); }

/* ===================================================================================== */
/* Hamiltonian terms                                                                     */
/* ===================================================================================== */
/* Diagonal element: stored/H_local.f90:13-71 == direct/HxV_local.f90:15-73, same order. */
static double h_local_element(const orc_ctx *c, const int *nup, const int *ndw) {
  const int norb = c->norb, nspin = c->nspin, nbath = c->nbath;
  const int sl = nspin - 1;                                /* Fortran index Nspin */
  double htmp = 0.0;
  for (int io = 0; io < norb; io++) {
    htmp = htmp + HLOC(c, 0, 0, io, io) * (double)nup[io];
    htmp = htmp + HLOC(c, sl, sl, io, io) * (double)ndw[io];
    htmp = htmp - c->xmu * (double)(nup[io] + ndw[io]);
  }
  for (int io = 0; io < norb; io++)
    htmp = htmp + c->uloc[io] * (double)nup[io] * (double)ndw[io];
  if (norb > 1) {
    for (int io = 0; io < norb; io++)
      for (int jo = io + 1; jo < norb; jo++)
        htmp = htmp + c->ust * (double)(nup[io] * ndw[jo] + nup[jo] * ndw[io]);
    for (int io = 0; io < norb; io++)
      for (int jo = io + 1; jo < norb; jo++)
        htmp = htmp + (c->ust - c->jh) * (double)(nup[io] * nup[jo] + ndw[io] * ndw[jo]);
  }
  if (c->hfmode) {
    for (int io = 0; io < norb; io++)
      htmp = htmp - 0.5 * c->uloc[io] * (double)(nup[io] + ndw[io]) + 0.25 * c->uloc[io];
    if (norb > 1) {
      for (int io = 0; io < norb; io++)
        for (int jo = io + 1; jo < norb; jo++) {
          htmp = htmp - 0.5 * c->ust * (double)(nup[io] + ndw[io] + nup[jo] + ndw[jo]) + 0.25 * c->ust;
          htmp = htmp - 0.5 * (c->ust - c->jh) * (double)(nup[io] + ndw[io] + nup[jo] + ndw[jo]) +
                 0.25 * (c->ust - c->jh);
        }
    }
  }
  for (int io = 0; io < c->nfoo; io++)                     /* size(bath_diag,2): Norb (normal, replica), 1 (hybrid) */
    for (int kp = 0; kp < nbath; kp++) {
      int ialfa = orc_bath_stride(c, io + 1, kp + 1) - 1;
      htmp = htmp + BATHE(c, 0, io, kp) * (double)nup[ialfa];
      htmp = htmp + BATHE(c, sl, io, kp) * (double)ndw[ialfa];
    }
  return htmp;
}

/* One-spin hopping terms.  Calls emit(target_state, value) for every hop out of source state
 * m, in the order of stored/H_up.f90:8-81 (== H_dw.f90, direct/HxV_up.f90, HxV_dw.f90).
 * is = spin index used for impHloc/diag_hybr (0 for up, Nspin-1 for dw). */
typedef void (*hop_emit_fn)(void *u, int32_t k2, double htmp);
static void one_spin_hops(const orc_ctx *c, int is, int32_t m, hop_emit_fn emit, void *u) {
  int n[64];
  bdecomp(m, c->ns, n);
  int32_t k1, k2; double sg1, sg2;
  for (int io = 0; io < c->norb; io++)
    for (int jo = 0; jo < c->norb; jo++) {
      if (HLOC(c, is, is, io, jo) != 0.0 && n[jo] == 1 && n[io] == 0) {
        orc_c(jo + 1, m, &k1, &sg1);
        orc_cdg(io + 1, k1, &k2, &sg2);
        emit(u, k2, HLOC(c, is, is, io, jo) * sg1 * sg2);
      }
    }
  if (c->bath_type == 2)                                   /* replica inter-orbital bath hopping, H_up.f90:26-50 */
    for (int kp = 0; kp < c->nbath; kp++)
      for (int io = 0; io < c->norb; io++)
        for (int jo = 0; jo < c->norb; jo++) {
          int ialfa = orc_bath_stride(c, io + 1, kp + 1), ibeta = orc_bath_stride(c, jo + 1, kp + 1);
          double hb = c->bath_h[is + c->nspin * (is + c->nspin * (io + c->norb * (jo + c->norb * (size_t)kp)))];
          if (hb != 0.0 && n[ibeta - 1] == 1 && n[ialfa - 1] == 0) {
            orc_c(ibeta, m, &k1, &sg1);
            orc_cdg(ialfa, k1, &k2, &sg2);
            emit(u, k2, hb * sg1 * sg2);
          }
        }
  for (int io = 0; io < c->norb; io++)
    for (int kp = 0; kp < c->nbath; kp++) {
      int ialfa = orc_bath_stride(c, io + 1, kp + 1);      /* 1-based */
      double vv = BATHV(c, is, io, kp);
      if (vv != 0.0 && n[io] == 1 && n[ialfa - 1] == 0) {
        orc_c(io + 1, m, &k1, &sg1);
        orc_cdg(ialfa, k1, &k2, &sg2);
        emit(u, k2, vv * sg1 * sg2);
      }
      if (vv != 0.0 && n[io] == 0 && n[ialfa - 1] == 1) {
        orc_c(ialfa, m, &k1, &sg1);
        orc_cdg(io + 1, k1, &k2, &sg2);
        emit(u, k2, vv * sg1 * sg2);
      }
    }
}

/* Non-local (spin-exchange, pair-hopping) terms out of (mup,mdw), in the order of
 * stored/H_non_local.f90:21-83 (== direct/HxV_non_local.f90:17-69). */
typedef void (*nl_emit_fn)(void *u, int32_t kup, int32_t kdw, double htmp);
static void non_local_hops(const orc_ctx *c, int32_t mup, int32_t mdw, nl_emit_fn emit, void *u) {
  int nup[64], ndw[64];
  bdecomp(mup, c->ns, nup);
  bdecomp(mdw, c->ns, ndw);
  int32_t k1, k2, k3, k4; double sg1, sg2, sg3, sg4;
  if (c->jhflag && c->jx != 0.0)
    for (int io = 0; io < c->norb; io++)
      for (int jo = 0; jo < c->norb; jo++)
        if (io != jo && nup[jo] == 1 && ndw[io] == 1 && ndw[jo] == 0 && nup[io] == 0) {
          orc_c(io + 1, mdw, &k1, &sg1);
          orc_cdg(jo + 1, k1, &k2, &sg2);
          orc_c(jo + 1, mup, &k3, &sg3);
          orc_cdg(io + 1, k3, &k4, &sg4);
          emit(u, k4, k2, c->jx * sg1 * sg2 * sg3 * sg4);
        }
  if (c->jhflag && c->jp != 0.0)
    for (int io = 0; io < c->norb; io++)
      for (int jo = 0; jo < c->norb; jo++)
        if (nup[jo] == 1 && ndw[jo] == 1 && ndw[io] == 0 && nup[io] == 0) {
          orc_c(jo + 1, mdw, &k1, &sg1);
          orc_cdg(io + 1, k1, &k2, &sg2);
          orc_c(jo + 1, mup, &k3, &sg3);
          orc_cdg(io + 1, k3, &k4, &sg4);
          emit(u, k4, k2, c->jp * sg1 * sg2 * sg3 * sg4);
        }
}

/* ===================================================================================== */
/* build_Hv_sector / ed_buildh_main                                                      */
/* ===================================================================================== */
typedef struct { rl_mat *m; const int32_t *map; int64_t n; int64_t j; } fac_emit_ctx;
static void fac_emit(void *u, int32_t k2, double htmp) {
  fac_emit_ctx *e = (fac_emit_ctx *)u;
  int64_t i = orc_binary_search(e->map, e->n, k2) - 1;
  rl_insert(e->m, htmp, i, e->j);                          /* (row = target, col = source) */
}
typedef struct { rl_mat *m; const orc_sector *s; int64_t irow; } nd_emit_ctx;
static void nd_emit(void *u, int32_t kup, int32_t kdw, double htmp) {
  nd_emit_ctx *e = (nd_emit_ctx *)u;
  int64_t jup = orc_binary_search(e->s->map_up, e->s->dimup, kup) - 1;
  int64_t jdw = orc_binary_search(e->s->map_dw, e->s->dimdw, kdw) - 1;
  rl_insert(e->m, htmp, e->irow, jup + jdw * e->s->dimup); /* global electron index column */
}

int64_t orc_vecdim_hv_sector(const orc_ctx *c, int nup, int ndw, int rank, int nranks) {
  /* vecDim_Hv_sector, ED_HAMILTONIAN.f90:229-253 (DimPh = 1) */
  int64_t dimup = orc_binomial(c->ns, nup), dimdw = orc_binomial(c->ns, ndw);
  int64_t q = dimdw / nranks;
  if (rank < dimdw % nranks) q++;
  return dimup * q;
}

orc_sector *orc_build_hv_sector(const orc_ctx *c, int nup, int ndw, int rank, int nranks,
                                int sparse_h) {
  orc_sector *s = (orc_sector *)xcalloc(1, sizeof(orc_sector));
  s->ctx = c; s->nup = nup; s->ndw = ndw; s->rank = rank; s->nranks = nranks;
  s->sparse_h = sparse_h;
  s->dimup = orc_binomial(c->ns, nup);
  s->dimdw = orc_binomial(c->ns, ndw);
  s->dim = s->dimup * s->dimdw;
  s->map_up = (int32_t *)xmalloc((size_t)s->dimup * sizeof(int32_t));
  s->map_dw = (int32_t *)xmalloc((size_t)s->dimdw * sizeof(int32_t));
  orc_build_sector_map(c->ns, nup, s->map_up);
  orc_build_sector_map(c->ns, ndw, s->map_dw);
  /* dw split, ED_HAMILTONIAN.f90:96-110 */
  s->qdw = s->dimdw / nranks;
  s->rdw = s->dimdw % nranks;
  if (rank < s->dimdw % nranks) { s->rdw = 0; s->qdw = s->qdw + 1; }
  s->q = s->dimup * s->qdw;
  s->r = s->dimup * s->rdw;
  s->istart = rank * s->q + s->r;            /* 0-based first row */
  s->iend = (rank + 1) * s->q + s->r;        /* one past last row */
  s->ishift = rank * s->q + s->r;
  if (!sparse_h) return s;

  /* ed_buildh_main, ED_HAMILTONIAN_SPARSE_HxV.f90:25-200 */
  int nupv[64], ndwv[64];
  /* stored/H_local.f90:1-80 */
  s->h0d = (double *)xmalloc((size_t)(s->iend - s->istart) * sizeof(double));
  for (int64_t i = s->istart; i < s->iend; i++) {
    int64_t iup = i % s->dimup, idw = i / s->dimup;
    bdecomp(s->map_up[iup], c->ns, nupv);
    bdecomp(s->map_dw[idw], c->ns, ndwv);
    s->h0d[i - s->ishift] = h_local_element(c, nupv, ndwv);
  }
  /* stored/H_non_local.f90:4-85 */
  if (c->jhflag) {
    rl_mat nd; rl_init(&nd, s->iend - s->istart); nd.ncol = s->dim;
    for (int64_t i = s->istart; i < s->iend; i++) {
      int64_t iup = i % s->dimup, idw = i / s->dimup;
      nd_emit_ctx e = { &nd, s, i - s->ishift };
      non_local_hops(c, s->map_up[iup], s->map_dw[idw], nd_emit, &e);
    }
    rl_to_csr(&nd, &s->hnd);
  }
  /* stored/H_up.f90:1-83 */
  {
    rl_mat up; rl_init(&up, s->dimup);
    for (int64_t jup = 0; jup < s->dimup; jup++) {
      fac_emit_ctx e = { &up, s->map_up, s->dimup, jup };
      one_spin_hops(c, 0, s->map_up[jup], fac_emit, &e);
    }
    rl_to_csr(&up, &s->hup);
  }
  /* stored/H_dw.f90:1-82 */
  {
    rl_mat dw; rl_init(&dw, s->dimdw);
    for (int64_t jdw = 0; jdw < s->dimdw; jdw++) {
      fac_emit_ctx e = { &dw, s->map_dw, s->dimdw, jdw };
      one_spin_hops(c, c->nspin - 1, s->map_dw[jdw], fac_emit, &e);
    }
    rl_to_csr(&dw, &s->hdw);
  }
  return s;
}

void orc_delete_hv_sector(orc_sector *s) {               /* ED_HAMILTONIAN.f90:174-222 */
  if (!s) return;
  free(s->map_up); free(s->map_dw); free(s->h0d);
  csr_free(&s->hup); csr_free(&s->hdw); csr_free(&s->hnd);
  free(s);
}

/* Dense Hmat = diag + Hnd + kron(Hdw, I_up) + kron(I_dw, Hup):
 * ED_HAMILTONIAN_SPARSE_HxV.f90:132-166.  Serial sectors only (rank 0 of 1). */
void orc_build_hmat(const orc_sector *s, double *hmat) {
  int64_t n = s->dim, du = s->dimup, dd = s->dimdw;
  memset(hmat, 0, (size_t)n * n * sizeof(double));
  for (int64_t i = 0; i < n; i++) hmat[i + n * i] += s->h0d[i];
  if (s->ctx->jhflag)
    for (int64_t i = 0; i < n; i++)
      for (int64_t p = s->hnd.rowptr[i]; p < s->hnd.rowptr[i + 1]; p++)
        hmat[i + n * s->hnd.cols[p]] += s->hnd.vals[p];
  for (int64_t idw = 0; idw < dd; idw++)
    for (int64_t p = s->hdw.rowptr[idw]; p < s->hdw.rowptr[idw + 1]; p++)
      for (int64_t iup = 0; iup < du; iup++)
        hmat[(iup + idw * du) + n * (iup + s->hdw.cols[p] * du)] += s->hdw.vals[p];
  for (int64_t idw = 0; idw < dd; idw++)
    for (int64_t iup = 0; iup < du; iup++)
      for (int64_t p = s->hup.rowptr[iup]; p < s->hup.rowptr[iup + 1]; p++)
        hmat[(iup + idw * du) + n * (s->hup.cols[p] + idw * du)] += s->hup.vals[p];
}

/* ===================================================================================== */
/* spMatVec_main, ED_HAMILTONIAN_SPARSE_HxV.f90:391-485 (DimPh = 1)                      */
/* ===================================================================================== */
void orc_spmatvec_main(const orc_sector *s, int64_t nloc, const double *v, double *hv) {
  const int64_t du = s->dimup, dd = s->dimdw;
  if (nloc != s->dim) { fprintf(stderr, "spMatVec_main: Nloc != dim\n"); abort(); }
  for (int64_t i = 0; i < nloc; i++) hv[i] = 0.0;
  /* Local */
  for (int64_t i = 0; i < nloc; i++) hv[i] = hv[i] + s->h0d[i] * v[i];
  /* DW: outer iup, inner idw (:414-427) */
  for (int64_t iup = 0; iup < du; iup++)
    for (int64_t idw = 0; idw < dd; idw++) {
      int64_t i = iup + idw * du;
      for (int64_t jj = s->hdw.rowptr[idw]; jj < s->hdw.rowptr[idw + 1]; jj++) {
        int64_t j = iup + s->hdw.cols[jj] * du;
        hv[i] = hv[i] + s->hdw.vals[jj] * v[j];
      }
    }
  /* UP (:430-443) */
  for (int64_t idw = 0; idw < dd; idw++)
    for (int64_t iup = 0; iup < du; iup++) {
      int64_t i = iup + idw * du;
      for (int64_t jj = s->hup.rowptr[iup]; jj < s->hup.rowptr[iup + 1]; jj++) {
        int64_t j = s->hup.cols[jj] + idw * du;
        hv[i] = hv[i] + s->hup.vals[jj] * v[j];
      }
    }
  /* Non-local (:473-483) */
  if (s->ctx->jhflag)
    for (int64_t i = 0; i < nloc; i++)
      for (int64_t jj = s->hnd.rowptr[i]; jj < s->hnd.rowptr[i + 1]; jj++)
        hv[i] = hv[i] + s->hnd.vals[jj] * v[s->hnd.cols[jj]];
}

/* spMatVec_main restricted to the i_dw column block owned by sector view s (rank r of P):
 * the same loop nests and term order as orc_spmatvec_main, rows istart..iend only, reading the
 * full vector.  Used by bench.py to time a BOUNDED sample of the CPU path (one block per thread
 * = the work one MPI rank of the reference would do, communication excluded). */
void orc_spmatvec_block(const orc_sector *s, const double *v_full, double *hv_block) {
  const int64_t du = s->dimup, c0 = s->istart / du, c1 = s->iend / du;
  const int64_t nloc = s->iend - s->istart, sh = s->ishift;
  for (int64_t i = 0; i < nloc; i++) hv_block[i] = 0.0;
  for (int64_t i = 0; i < nloc; i++) hv_block[i] = hv_block[i] + s->h0d[i] * v_full[i + sh];
  for (int64_t iup = 0; iup < du; iup++)
    for (int64_t idw = c0; idw < c1; idw++) {
      int64_t i = iup + idw * du - sh;
      for (int64_t jj = s->hdw.rowptr[idw]; jj < s->hdw.rowptr[idw + 1]; jj++)
        hv_block[i] = hv_block[i] + s->hdw.vals[jj] * v_full[iup + s->hdw.cols[jj] * du];
    }
  for (int64_t idw = c0; idw < c1; idw++)
    for (int64_t iup = 0; iup < du; iup++) {
      int64_t i = iup + idw * du - sh;
      for (int64_t jj = s->hup.rowptr[iup]; jj < s->hup.rowptr[iup + 1]; jj++)
        hv_block[i] = hv_block[i] + s->hup.vals[jj] * v_full[s->hup.cols[jj] + idw * du];
    }
  if (s->ctx->jhflag)
    for (int64_t i = 0; i < nloc; i++)
      for (int64_t jj = s->hnd.rowptr[i]; jj < s->hnd.rowptr[i + 1]; jj++)
        hv_block[i] = hv_block[i] + s->hnd.vals[jj] * v_full[s->hnd.cols[jj]];
}
/* The same loops as orc_spmatvec_block (same summation order, same results) for a caller that holds only the
 * columns of v the block touches: colidx[0..ncolidx) = ascending global i_dw indices, xcols = those columns, DimUp
 * values each, in that order.  Used for spot checks of sectors whose full vector does not fit the host
 * (Ns = 18: 19 GB).  Returns 0, or 1 if a needed column is missing. */
static int64_t col_slot(const int64_t *colidx, int64_t n, int64_t col) {
  int64_t lo = 0, hi = n - 1;
  while (lo <= hi) {
    int64_t mid = (lo + hi) / 2;
    if (colidx[mid] == col) return mid;
    if (colidx[mid] < col) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}
int orc_spmatvec_block_cols(const orc_sector *s, const int64_t *colidx, int64_t ncolidx, const double *xcols,
                            double *hv_block) {
  const int64_t du = s->dimup, c0 = s->istart / du, c1 = s->iend / du;
  const int64_t nloc = s->iend - s->istart, sh = s->ishift;
  if (s->ctx->jhflag) return 1;                       /* spH0nd needs the whole vector */
  for (int64_t i = 0; i < nloc; i++) hv_block[i] = 0.0;
  for (int64_t idw = c0; idw < c1; idw++) {
    int64_t p = col_slot(colidx, ncolidx, idw);
    if (p < 0) return 1;
    for (int64_t iup = 0; iup < du; iup++) {
      int64_t i = iup + idw * du - sh;
      hv_block[i] = hv_block[i] + s->h0d[i] * xcols[iup + p * du];
    }
  }
  for (int64_t iup = 0; iup < du; iup++)
    for (int64_t idw = c0; idw < c1; idw++) {
      int64_t i = iup + idw * du - sh;
      for (int64_t jj = s->hdw.rowptr[idw]; jj < s->hdw.rowptr[idw + 1]; jj++) {
        int64_t p = col_slot(colidx, ncolidx, s->hdw.cols[jj]);
        if (p < 0) return 1;
        hv_block[i] = hv_block[i] + s->hdw.vals[jj] * xcols[iup + p * du];
      }
    }
  for (int64_t idw = c0; idw < c1; idw++) {
    int64_t p = col_slot(colidx, ncolidx, idw);
    for (int64_t iup = 0; iup < du; iup++) {
      int64_t i = iup + idw * du - sh;
      for (int64_t jj = s->hup.rowptr[iup]; jj < s->hup.rowptr[iup + 1]; jj++)
        hv_block[i] = hv_block[i] + s->hup.vals[jj] * xcols[s->hup.cols[jj] + p * du];
    }
  }
  return 0;
}
typedef struct { orc_sector **secs; const double *v; double **out; } blk_job;
static void blk_rank(int r, void *arg) {
  blk_job *b = (blk_job *)arg;
  orc_spmatvec_block(b->secs[r], b->v, b->out[r]);
}
/* nblk blocks on nthreads threads; returns the number of vector elements processed */
int64_t orc_spmatvec_blocks_mt(orc_sector **secs, int nblk, int nthreads, const double *v_full,
                               double **hv_blocks) {
  blk_job b = { secs, v_full, hv_blocks };
  par_for(nthreads, nblk, blk_rank, &b);
  int64_t n = 0;
  for (int r = 0; r < nblk; r++) n += secs[r]->iend - secs[r]->istart;
  return n;
}

/* ===================================================================================== */
/* directMatVec_main, ED_HAMILTONIAN_DIRECT_HxV.f90:21-95 + direct/HxV_*.f90 (scatter)   */
/* ===================================================================================== */
typedef struct {
  const int32_t *map; int64_t n;    /* factor map */
  double *hv; double vin;           /* target vector and source value */
  int64_t stride, other;            /* i = pos*stride + other */
} dir_emit_ctx;
static void dir_emit(void *u, int32_t k2, double htmp) {
  dir_emit_ctx *e = (dir_emit_ctx *)u;
  int64_t p = orc_binary_search(e->map, e->n, k2) - 1;
  int64_t i = p * e->stride + e->other;
  e->hv[i] = e->hv[i] + htmp * e->vin;
}
typedef struct { const orc_sector *s; double *hv; double vin; } dirnl_emit_ctx;
static void dirnl_emit(void *u, int32_t kup, int32_t kdw, double htmp) {
  dirnl_emit_ctx *e = (dirnl_emit_ctx *)u;
  int64_t iup = orc_binary_search(e->s->map_up, e->s->dimup, kup) - 1;
  int64_t idw = orc_binary_search(e->s->map_dw, e->s->dimdw, kdw) - 1;
  int64_t i = iup + idw * e->s->dimup;
  e->hv[i] = e->hv[i] + htmp * e->vin;
}

void orc_directmatvec_main(const orc_sector *s, int64_t nloc, const double *vin, double *hv) {
  const orc_ctx *c = s->ctx;
  const int64_t du = s->dimup, dd = s->dimdw;
  if (nloc != s->dim) { fprintf(stderr, "directMatVec_cc ERROR: Nloc != dim(isector)\n"); abort(); }
  int nupv[64], ndwv[64];
  for (int64_t i = 0; i < nloc; i++) hv[i] = 0.0;
  /* direct/HxV_local.f90 */
  for (int64_t i = 0; i < nloc; i++) {
    bdecomp(s->map_up[i % du], c->ns, nupv);
    bdecomp(s->map_dw[i / du], c->ns, ndwv);
    double htmp = h_local_element(c, nupv, ndwv);
    hv[i] = hv[i] + htmp * vin[i];
  }
  /* direct/HxV_up.f90: loop jdw, jup */
  for (int64_t jdw = 0; jdw < dd; jdw++)
    for (int64_t jup = 0; jup < du; jup++) {
      dir_emit_ctx e = { s->map_up, du, hv, vin[jup + jdw * du], 1, jdw * du };
      one_spin_hops(c, 0, s->map_up[jup], dir_emit, &e);
    }
  /* direct/HxV_dw.f90: loop jup, jdw */
  for (int64_t jup = 0; jup < du; jup++)
    for (int64_t jdw = 0; jdw < dd; jdw++) {
      dir_emit_ctx e = { s->map_dw, dd, hv, vin[jup + jdw * du], du, jup };
      one_spin_hops(c, c->nspin - 1, s->map_dw[jdw], dir_emit, &e);
    }
  /* direct/HxV_non_local.f90 */
  if (c->jhflag)
    for (int64_t j = 0; j < nloc; j++) {
      dirnl_emit_ctx e = { s, hv, vin[j] };
      non_local_hops(c, s->map_up[j % du], s->map_dw[j / du], dirnl_emit, &e);
    }
}

/* ===================================================================================== */
/* vector_transpose_MPI for all ranks, ED_HAMILTONIAN_COMMON.f90:53-125                  */
/* a_shards[r]: nrow x qcol(r) column-major; b_shards[r]: ncol x qrow(r) column-major.   */
/* ===================================================================================== */
static int64_t split_q(int64_t n, int p, int r) { return n / p + ((r < n % p) ? 1 : 0); }
static int64_t split_off(int64_t n, int p, int r) {
  int64_t q = n / p, m = n % p;
  return (r < m) ? r * (q + 1) : m * (q + 1) + (r - m) * q;
}
typedef struct { int nranks; int64_t nrow, ncol; double *const *a; double *const *b; } tr_job;
static void transpose_recv_rank(int d, void *arg) {
  /* what destination rank d ends up with after the Ntot per-column MPI_AllToAllV calls
   * (:107-113) followed by local_transpose (:115,121-125) */
  tr_job *t = (tr_job *)arg;
  int P = t->nranks;
  int64_t nrow = t->nrow, ncol = t->ncol;
  int64_t qr = split_q(nrow, P, d), ro = split_off(nrow, P, d);
  int64_t ntot = ncol / P + ((ncol % P) ? 1 : 0);
  double *tmp = (double *)xcalloc((size_t)(qr * ncol), sizeof(double));
  for (int64_t j = 0; j < ntot; j++)
    for (int s = 0; s < P; s++) {
      if (j >= split_q(ncol, P, s)) continue;              /* zero send counts */
      const double *col = t->a[s] + j * nrow;
      int64_t g = split_off(ncol, P, s) + j;               /* global column index */
      memcpy(tmp + g * qr, col + ro, (size_t)qr * sizeof(double)); /* recv_offset = g*qrow(d) */
    }
  /* local_transpose: mat(ncol,qrow) = transpose(reshape(mat,[qrow,ncol])) */
  for (int64_t r = 0; r < qr; r++)
    for (int64_t g = 0; g < ncol; g++) t->b[d][g + r * ncol] = tmp[r + g * qr];
  free(tmp);
}
static void transpose_all_mt(int nranks, int nthreads, int64_t nrow, int64_t ncol,
                             double *const *a, double *const *b) {
  tr_job t = { nranks, nrow, ncol, a, b };
  par_for(nthreads, nranks, transpose_recv_rank, &t);
}
void orc_vector_transpose_all(int nranks, int64_t nrow, int64_t ncol,
                              double *const *a, double *const *b) {
  transpose_all_mt(nranks, 1, nrow, ncol, a, b);
}

/* ===================================================================================== */
/* spMatVec_mpi_main for all ranks, ED_HAMILTONIAN_SPARSE_HxV.f90:568-694                */
/* ===================================================================================== */
typedef struct {
  orc_sector **secs; int P; const orc_ctx *c;
  const double *v; double **vloc, **hloc, **vt, **hvt, **back;
} mpi_job;

static void sp_rank_local_up(int r, void *arg) {           /* :587-613 */
  mpi_job *J = (mpi_job *)arg;
  const orc_sector *s = J->secs[r];
  const int64_t du = s->dimup, nloc = du * s->qdw;
  double *h = J->hloc[r]; const double *x = J->vloc[r];
  for (int64_t i = 0; i < nloc; i++) h[i] = 0.0;
  for (int64_t i = 0; i < nloc; i++) h[i] = h[i] + s->h0d[i] * x[i];
  for (int64_t idw = 0; idw < s->qdw; idw++)
    for (int64_t iup = 0; iup < du; iup++) {
      int64_t i = iup + idw * du;
      for (int64_t jj = s->hup.rowptr[iup]; jj < s->hup.rowptr[iup + 1]; jj++)
        h[i] = h[i] + s->hup.vals[jj] * x[s->hup.cols[jj] + idw * du];
    }
}
static void sp_rank_dw(int r, void *arg) {                 /* :628-639, names swapped there */
  mpi_job *J = (mpi_job *)arg;
  const orc_sector *s = J->secs[r];
  const int64_t dd = s->dimdw, qup = split_q(s->dimup, J->P, r);
  double *hvt = J->hvt[r]; const double *vt = J->vt[r];
  for (int64_t idw = 0; idw < qup; idw++)
    for (int64_t iup = 0; iup < dd; iup++) {
      int64_t i = iup + idw * dd;
      for (int64_t jj = s->hdw.rowptr[iup]; jj < s->hdw.rowptr[iup + 1]; jj++)
        hvt[i] = hvt[i] + s->hdw.vals[jj] * vt[s->hdw.cols[jj] + idw * dd];
    }
}
static void rank_add_back(int r, void *arg) {              /* :642 */
  mpi_job *J = (mpi_job *)arg;
  int64_t nloc = J->secs[r]->dimup * J->secs[r]->qdw;
  for (int64_t i = 0; i < nloc; i++) J->hloc[r][i] = J->hloc[r][i] + J->back[r][i];
}
static void sp_rank_nonlocal(int r, void *arg) {           /* :673-692; allgather(v) == v */
  mpi_job *J = (mpi_job *)arg;
  const orc_sector *s = J->secs[r];
  int64_t nloc = s->dimup * s->qdw;
  for (int64_t i = 0; i < nloc; i++)
    for (int64_t jj = s->hnd.rowptr[i]; jj < s->hnd.rowptr[i + 1]; jj++)
      J->hloc[r][i] = J->hloc[r][i] + s->hnd.vals[jj] * J->v[s->hnd.cols[jj]];
}

static void mpi_job_alloc(mpi_job *J, orc_sector **secs, int P, const double *v, double *hv) {
  const int64_t du = secs[0]->dimup, dd = secs[0]->dimdw;
  J->secs = secs; J->P = P; J->c = secs[0]->ctx; J->v = v;
  J->vloc = (double **)xmalloc((size_t)P * sizeof(double *));
  J->hloc = (double **)xmalloc((size_t)P * sizeof(double *));
  J->vt = (double **)xmalloc((size_t)P * sizeof(double *));
  J->hvt = (double **)xmalloc((size_t)P * sizeof(double *));
  J->back = (double **)xmalloc((size_t)P * sizeof(double *));
  for (int r = 0; r < P; r++) {
    J->vloc[r] = (double *)v + secs[r]->ishift;
    J->hloc[r] = hv + secs[r]->ishift;
    int64_t qup = split_q(du, P, r);
    /* the reference allocates and zeroes vt/Hvt on every call (:621-625) */
    J->vt[r] = (double *)xcalloc((size_t)(qup * dd), sizeof(double));
    J->hvt[r] = (double *)xcalloc((size_t)(qup * dd), sizeof(double));
    J->back[r] = (double *)xcalloc((size_t)(du * secs[r]->qdw), sizeof(double));
  }
}
static void mpi_job_free(mpi_job *J) {
  for (int r = 0; r < J->P; r++) { free(J->vt[r]); free(J->hvt[r]); free(J->back[r]); }
  free(J->vloc); free(J->hloc); free(J->vt); free(J->hvt); free(J->back);
}

void orc_spmatvec_mpi_main_prebuilt(orc_sector **secs, int P, int nthreads,
                                    const double *v, double *hv) {
  mpi_job J;
  mpi_job_alloc(&J, secs, P, v, hv);
  const int64_t du = secs[0]->dimup, dd = secs[0]->dimdw;
  par_for(nthreads, P, sp_rank_local_up, &J);
  transpose_all_mt(P, nthreads, du, dd, J.vloc, J.vt);
  par_for(nthreads, P, sp_rank_dw, &J);
  transpose_all_mt(P, nthreads, dd, du, J.hvt, J.back);
  par_for(nthreads, P, rank_add_back, &J);
  if (J.c->jhflag) par_for(nthreads, P, sp_rank_nonlocal, &J);
  mpi_job_free(&J);
}

void orc_spmatvec_mpi_main_all(const orc_ctx *c, int nup, int ndw, int P, int nthreads,
                               const double *v, double *hv) {
  orc_sector **secs = (orc_sector **)xmalloc((size_t)P * sizeof(orc_sector *));
  for (int r = 0; r < P; r++) secs[r] = orc_build_hv_sector(c, nup, ndw, r, P, 1);
  orc_spmatvec_mpi_main_prebuilt(secs, P, nthreads, v, hv);
  for (int r = 0; r < P; r++) orc_delete_hv_sector(secs[r]);
  free(secs);
}

/* directMatVec_MPI_main for all ranks, ED_HAMILTONIAN_DIRECT_HxV.f90:180-284 +
 * direct_mpi/HxV_*.f90 */
typedef struct { const orc_sector *s; const double *vt; double *hvj; } dirnlg_emit_ctx;
static void dirnlg_emit(void *u, int32_t kup, int32_t kdw, double htmp) {
  dirnlg_emit_ctx *e = (dirnlg_emit_ctx *)u;            /* gather form, HxV_non_local.f90:36 */
  int64_t iup = orc_binary_search(e->s->map_up, e->s->dimup, kup) - 1;
  int64_t idw = orc_binary_search(e->s->map_dw, e->s->dimdw, kdw) - 1;
  *e->hvj = *e->hvj + htmp * e->vt[iup + idw * e->s->dimup];
}
static void dir_rank_local_up(int r, void *arg) {
  mpi_job *J = (mpi_job *)arg;
  const orc_sector *s = J->secs[r];
  const orc_ctx *c = J->c;
  const int64_t du = s->dimup, nloc = du * s->qdw;
  int nupv[64], ndwv[64];
  for (int64_t i = 0; i < nloc; i++) J->hloc[r][i] = 0.0;
  for (int64_t i = 0; i < nloc; i++) {                    /* direct_mpi/HxV_local.f90 */
    int64_t ig = i + s->ishift;
    bdecomp(s->map_up[ig % du], c->ns, nupv);
    bdecomp(s->map_dw[ig / du], c->ns, ndwv);
    J->hloc[r][i] = J->hloc[r][i] + h_local_element(c, nupv, ndwv) * J->vloc[r][i];
  }
  for (int64_t jdw = 0; jdw < s->qdw; jdw++)              /* direct_mpi/HxV_up.f90 */
    for (int64_t jup = 0; jup < du; jup++) {
      dir_emit_ctx e = { s->map_up, du, J->hloc[r], J->vloc[r][jup + jdw * du], 1, jdw * du };
      one_spin_hops(c, 0, s->map_up[jup], dir_emit, &e);
    }
}
static void dir_rank_dw(int r, void *arg) {               /* direct_mpi/HxV_dw.f90 */
  mpi_job *J = (mpi_job *)arg;
  const orc_sector *s = J->secs[r];
  const orc_ctx *c = J->c;
  const int64_t dd = s->dimdw, qup = split_q(s->dimup, J->P, r);
  for (int64_t jdw = 0; jdw < qup; jdw++)
    for (int64_t jup = 0; jup < dd; jup++) {
      dir_emit_ctx e = { s->map_dw, dd, J->hvt[r], J->vt[r][jup + jdw * dd], 1, jdw * dd };
      one_spin_hops(c, c->nspin - 1, s->map_dw[jup], dir_emit, &e);
    }
}
static void dir_rank_nonlocal(int r, void *arg) {         /* direct_mpi/HxV_non_local.f90 */
  mpi_job *J = (mpi_job *)arg;
  const orc_sector *s = J->secs[r];
  const int64_t du = s->dimup, nloc = du * s->qdw;
  for (int64_t j = 0; j < nloc; j++) {
    int64_t jg = j + s->ishift;
    dirnlg_emit_ctx e = { s, J->v, &J->hloc[r][j] };
    non_local_hops(J->c, s->map_up[jg % du], s->map_dw[jg / du], dirnlg_emit, &e);
  }
}
void orc_directmatvec_mpi_main_all(const orc_ctx *c, int nup, int ndw, int P, int nthreads,
                                   const double *v, double *hv) {
  orc_sector **secs = (orc_sector **)xmalloc((size_t)P * sizeof(orc_sector *));
  for (int r = 0; r < P; r++) secs[r] = orc_build_hv_sector(c, nup, ndw, r, P, 0);
  mpi_job J;
  mpi_job_alloc(&J, secs, P, v, hv);
  const int64_t du = secs[0]->dimup, dd = secs[0]->dimdw;
  par_for(nthreads, P, dir_rank_local_up, &J);
  transpose_all_mt(P, nthreads, du, dd, J.vloc, J.vt);
  par_for(nthreads, P, dir_rank_dw, &J);
  transpose_all_mt(P, nthreads, dd, du, J.hvt, J.back);
  par_for(nthreads, P, rank_add_back, &J);
  if (c->jhflag) par_for(nthreads, P, dir_rank_nonlocal, &J);
  mpi_job_free(&J);
  for (int r = 0; r < P; r++) orc_delete_hv_sector(secs[r]);
  free(secs);
}

/* ===================================================================================== */
/* SciFortran SF_SP_LINALG simple Lanczos (restated; un-vendored dependency)             */
/* ===================================================================================== */
static double ddot(int64_t n, const double *a, const double *b) {
  double s = 0.0;
  for (int64_t i = 0; i < n; i++) s += a[i] * b[i];
  return s;
}

/* Symmetric tridiagonal QL with implicit shifts (EISPACK tql2 algorithm).  d[0..n-1]
 * diagonal, e[1..n-1] sub-diagonal (e[0] unused), z in: identity (column-major n x n),
 * out: eigenvectors; eigenvalues ascending in d. */
int orc_tql2(int n, double *d, double *e, double *z) {
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
        double c = 1.0, c2 = c, c3 = c, el1 = e[l1], s = 0.0, s2 = 0.0;
        for (int i = m - 1; i >= l; i--) {
          c3 = c2; c2 = c; s2 = s;
          g = c * e[i];
          h = c * p;
          r = hypot(p, e[i]);
          e[i + 1] = s * r;
          s = e[i] / r;
          c = p / r;
          p = c * d[i] - s * g;
          d[i + 1] = h + s * (c * g + s * d[i]);
          for (int k = 0; k < n; k++) {
            h = z[k + (size_t)n * (i + 1)];
            z[k + (size_t)n * (i + 1)] = s * z[k + (size_t)n * i] + c * h;
            z[k + (size_t)n * i] = c * z[k + (size_t)n * i] - s * h;
          }
        }
        p = -s * s2 * c3 * el1 * e[l] / dl1;
        e[l] = s * p;
        d[l] = c * p;
        tst2 = tst1 + fabs(e[l]);
      } while (tst2 > tst1);
    }
    d[l] += f;
  }
  for (int ii = 1; ii < n; ii++) {                          /* order ascending */
    int i = ii - 1, k = i;
    double p = d[i];
    for (int j = ii; j < n; j++)
      if (d[j] < p) { k = j; p = d[j]; }
    if (k != i) {
      d[k] = d[i]; d[i] = p;
      for (int j = 0; j < n; j++) {
        double t = z[j + (size_t)n * i];
        z[j + (size_t)n * i] = z[j + (size_t)n * k];
        z[j + (size_t)n * k] = t;
      }
    }
  }
  return 0;
}

/* lanczos_iteration: iter==1: v/=|v|; else t=v, v=w/beta, w=-beta*t;  w+=H v; a=v.w;
 * w-=a v; b=|w|.  (SciFortran SF_SP_LINALG, as used via ED_DIAG.f90:177, ED_GF_NORMAL.f90:232) */
static void lanczos_iteration(orc_matvec_fn mv, void *user, int64_t n, int iter,
                              double *vin, double *vout, double *tmp, double *alfa, double *beta) {
  if (iter == 1) {
    double norm = sqrt(ddot(n, vin, vin));
    if (norm == 0.0) { fprintf(stderr, "LANCZOS_ITERATION: norm = 0\n"); abort(); }
    for (int64_t i = 0; i < n; i++) vin[i] = vin[i] / norm;
  } else {
    double b = *beta;
    for (int64_t i = 0; i < n; i++) {
      double t = vin[i];
      vin[i] = vout[i] / b;
      vout[i] = -b * t;
    }
  }
  mv(user, n, vin, tmp);
  for (int64_t i = 0; i < n; i++) vout[i] = vout[i] + tmp[i];
  double a = ddot(n, vin, vout);
  for (int64_t i = 0; i < n; i++) vout[i] = vout[i] - a * vin[i];
  *alfa = a;
  *beta = sqrt(ddot(n, vout, vout));
}

static void tridiag_eig(int nlanc, const double *alanc, const double *blanc, double *diag, double *z) {
  double *sub = (double *)xcalloc((size_t)nlanc + 1, sizeof(double));
  for (int i = 0; i < nlanc; i++) diag[i] = alanc[i];
  for (int i = 1; i < nlanc; i++) sub[i] = blanc[i];
  memset(z, 0, (size_t)nlanc * nlanc * sizeof(double));
  for (int i = 0; i < nlanc; i++) z[i + (size_t)nlanc * i] = 1.0;
  if (orc_tql2(nlanc, diag, sub, z)) { fprintf(stderr, "tql2 failed\n"); abort(); }
  free(sub);
}

/* sp_lanc_eigh.  A zero start vector is replaced by a pseudo-random one (the reference uses
 * the compiler's random_number with a fixed seed, which is not reproducible across
 * compilers; here a fixed LCG -- pass an explicit start vector for parity runs). */
int orc_sp_lanc_eigh(orc_matvec_fn mv, void *user, int64_t n, double *egs, double *vect,
                     int nitermax, double threshold, int ncheck,
                     int *nlanc_out, double *alanc_out, double *blanc_out) {
  if (ncheck <= 0) ncheck = 10;
  double norm = ddot(n, vect, vect);
  if (norm == 0.0) {
    uint64_t st = 1234567ULL;
    for (int64_t i = 0; i < n; i++) {
      st = st * 6364136223846793005ULL + 1442695040888963407ULL;
      vect[i] = (double)(st >> 11) * (1.0 / 9007199254740992.0);
    }
    double nn = sqrt(ddot(n, vect, vect));
    for (int64_t i = 0; i < n; i++) vect[i] /= nn;
  }
  double *vin = (double *)xmalloc((size_t)n * sizeof(double));
  double *vout = (double *)xcalloc((size_t)n, sizeof(double));
  double *tmp = (double *)xmalloc((size_t)n * sizeof(double));
  double *alanc = (double *)xcalloc((size_t)nitermax + 2, sizeof(double));
  double *blanc = (double *)xcalloc((size_t)nitermax + 2, sizeof(double));
  double *diag = (double *)xcalloc((size_t)nitermax + 1, sizeof(double));
  double *z = (double *)xmalloc((size_t)nitermax * nitermax * sizeof(double));
  memcpy(vin, vect, (size_t)n * sizeof(double));
  int nlanc = 0;
  double a_ = 0.0, b_ = 0.0, esave = 0.0;
  *egs = 0.0;
  for (int iter = 1; iter <= nitermax; iter++) {
    lanczos_iteration(mv, user, n, iter, vin, vout, tmp, &a_, &b_);
    if (fabs(b_) < threshold) break;
    nlanc = nlanc + 1;
    alanc[iter - 1] = a_;
    blanc[iter] = b_;
    tridiag_eig(nlanc, alanc, blanc, diag, z);
    if (nlanc >= ncheck) {
      esave = diag[0];
      double diff = *egs - esave;
      *egs = esave;
      if (nlanc > ncheck && fabs(diff) <= threshold) break;
    }
  }
  if (nlanc == 0) { nlanc = 1; alanc[0] = a_; }            /* 1-step invariant subspace */
  tridiag_eig(nlanc, alanc, blanc, diag, z);
  *egs = diag[0];
  /* second sweep: rebuild the eigenvector  vect = sum_k Z(k,1) v_k */
  memcpy(vin, vect, (size_t)n * sizeof(double));
  for (int64_t i = 0; i < n; i++) { vout[i] = 0.0; vect[i] = 0.0; }
  double bprev = 0.0, aa;
  for (int iter = 1; iter <= nlanc; iter++) {
    lanczos_iteration(mv, user, n, iter, vin, vout, tmp, &aa, &bprev);
    double zk = z[(iter - 1) + 0 * (size_t)nlanc];
    for (int64_t i = 0; i < n; i++) vect[i] = vect[i] + vin[i] * zk;
  }
  norm = sqrt(ddot(n, vect, vect));
  for (int64_t i = 0; i < n; i++) vect[i] = vect[i] / norm;
  if (nlanc_out) *nlanc_out = nlanc;
  if (alanc_out) memcpy(alanc_out, alanc, (size_t)nlanc * sizeof(double));
  if (blanc_out) memcpy(blanc_out, blanc, (size_t)nlanc * sizeof(double));
  free(vin); free(vout); free(tmp); free(alanc); free(blanc); free(diag); free(z);
  return 0;
}

/* sp_lanc_tridiag: alanc(k)=a_k, blanc(k+1)=b_k, blanc(1)=0; exit when |b|<threshold.
 * Returns the number of completed steps.  vin is destroyed. */
int orc_sp_lanc_tridiag(orc_matvec_fn mv, void *user, int64_t n, double *vin,
                        double *alanc, double *blanc, int nitermax, double threshold) {
  double *vout = (double *)xcalloc((size_t)n, sizeof(double));
  double *tmp = (double *)xmalloc((size_t)n * sizeof(double));
  double a_ = 0.0, b_ = 0.0;
  int done = 0;
  for (int iter = 1; iter <= nitermax; iter++) {
    lanczos_iteration(mv, user, n, iter, vin, vout, tmp, &a_, &b_);
    alanc[iter - 1] = a_;
    done = iter;
    if (fabs(b_) < threshold) break;
    if (iter < nitermax) blanc[iter] = b_;
  }
  free(vout); free(tmp);
  return done;
}

typedef struct { const orc_sector *s; int mode; } sec_mv;
static void sec_matvec(void *u, int64_t n, const double *v, double *hv) {
  sec_mv *m = (sec_mv *)u;
  if (m->mode == 0) orc_spmatvec_main(m->s, n, v, hv);
  else orc_directmatvec_main(m->s, n, v, hv);
}
int orc_lanc_eigh_sector(const orc_sector *s, int mode, double *egs, double *vect,
                         int nitermax, double threshold, int ncheck,
                         int *nlanc_out, double *alanc_out, double *blanc_out) {
  sec_mv m = { s, mode };
  return orc_sp_lanc_eigh(sec_matvec, &m, s->dim, egs, vect, nitermax, threshold, ncheck,
                          nlanc_out, alanc_out, blanc_out);
}
int orc_lanc_tridiag_sector(const orc_sector *s, int mode, double *vin,
                            double *alanc, double *blanc, int nitermax, double threshold) {
  sec_mv m = { s, mode };
  return orc_sp_lanc_tridiag(sec_matvec, &m, s->dim, vin, alanc, blanc, nitermax, threshold);
}

/* ===================================================================================== */
/* ED_OBSERVABLES: lanc_observables (:95-363) + lanc_local_energy (:372-600), one state   */
/* ===================================================================================== */
void orc_observables_normal(const orc_ctx *c, int nup_, int ndw_, const double *gs, double zeta, orc_observables *o) {
  const int norb = c->norb, ns = c->ns, LD = ORC_MAX_ORB;
  const int64_t du = orc_binomial(ns, nup_), dd = orc_binomial(ns, ndw_), dim = du * dd;
  int32_t *hu = (int32_t *)xmalloc((size_t)du * 4), *hd = (int32_t *)xmalloc((size_t)dd * 4);
  orc_build_sector_map(ns, nup_, hu);
  orc_build_sector_map(ns, ndw_, hd);
  memset(o, 0, sizeof(*o));
  const double peso = 1.0 / zeta;                            /* T = 0: peso = 1/zeta_function (:144-145) */
  int ibup[64], ibdw[64];
  double nup[ORC_MAX_ORB], ndw[ORC_MAX_ORB], sz[ORC_MAX_ORB], nt[ORC_MAX_ORB];
  /* ---- first loop of lanc_observables (:153-212) ---- */
  for (int64_t i = 0; i < dim; i++) {
    int64_t iup = i % du, idw = i / du;
    int32_t mup = hu[iup], mdw = hd[idw];
    bdecomp(mup, ns, ibup);
    bdecomp(mdw, ns, ibdw);
    double gs_weight = peso * fabs(gs[i]) * fabs(gs[i]);
    for (int io = 0; io < norb; io++) {
      nup[io] = ibup[io]; ndw[io] = ibdw[io];
      sz[io] = (nup[io] - ndw[io]) / 2.0;
      nt[io] = nup[io] + ndw[io];
    }
    int iprob = 1, p3 = 1;
    for (int io = 0; io < norb; io++) { iprob += (int)lround(nt[io]) * p3; p3 *= 3; }
    o->prob[iprob - 1] += gs_weight;
    double ssum = 0.0;
    for (int io = 0; io < norb; io++) {
      o->dens[io] += nt[io] * gs_weight;
      o->dens_up[io] += nup[io] * gs_weight;
      o->dens_dw[io] += ndw[io] * gs_weight;
      o->docc[io] += nup[io] * ndw[io] * gs_weight;
      o->magz[io] += (nup[io] - ndw[io]) * gs_weight;
      o->sz2[io + LD * io] += (sz[io] * sz[io]) * gs_weight;
      o->n2[io + LD * io] += (nt[io] * nt[io]) * gs_weight;
      for (int jo = io + 1; jo < norb; jo++) {
        o->sz2[io + LD * jo] += (sz[io] * sz[jo]) * gs_weight;
        o->sz2[jo + LD * io] += (sz[jo] * sz[io]) * gs_weight;
        o->n2[io + LD * jo] += (nt[io] * nt[jo]) * gs_weight;
        o->n2[jo + LD * io] += (nt[jo] * nt[io]) * gs_weight;
      }
      ssum += sz[io];
    }
    o->s2tot += ssum * ssum * gs_weight;
  }
  /* ---- impurity density matrix (:240-305) ---- */
  for (int64_t i = 0; i < dim; i++) {
    int64_t iup = i % du, idw = i / du;
    int32_t iud[2] = { hu[iup], hd[idw] };
    bdecomp(iud[0], ns, ibup);
    bdecomp(iud[1], ns, ibdw);
    const int *nud[2] = { ibup, ibdw };
    for (int is = 0; is < c->nspin; is++)
      for (int io = 0; io < norb; io++)
        o->dm[is][io + LD * io] += peso * nud[is][io] * gs[i] * gs[i];
    for (int is = 0; is < c->nspin; is++)
      for (int io = 0; io < norb; io++)
        for (int jo = 0; jo < norb; jo++)
          if (nud[is][jo] == 1 && nud[is][io] == 0) {
            int32_t r, k; double sgn1, sgn2;
            orc_c(jo + 1, iud[is], &r, &sgn1);
            orc_cdg(io + 1, r, &k, &sgn2);
            int64_t ju = iup, jd = idw;
            if (is == 0) ju = orc_binary_search(hu, du, k) - 1;
            else         jd = orc_binary_search(hd, dd, k) - 1;
            int64_t j = ju + jd * du;
            o->dm[is][io + LD * jo] += peso * sgn1 * gs[i] * sgn2 * gs[j];
          }
  }
  /* ---- lanc_local_energy (:419-560) ---- */
  const int sl = c->nspin - 1;
  for (int64_t i = 0; i < dim; i++) {
    int64_t iup = i % du, idw = i / du;
    int32_t mup = hu[iup], mdw = hd[idw];
    bdecomp(mup, ns, ibup);
    bdecomp(mdw, ns, ibdw);
    for (int io = 0; io < norb; io++) { nup[io] = ibup[io]; ndw[io] = ibdw[io]; }
    double gs_weight = peso * fabs(gs[i]) * fabs(gs[i]);
    for (int io = 0; io < norb; io++) {
      o->eknot += HLOC(c, 0, 0, io, io) * nup[io] * gs_weight;
      o->eknot += HLOC(c, sl, sl, io, io) * ndw[io] * gs_weight;
    }
    for (int io = 0; io < norb; io++)
      for (int jo = 0; jo < norb; jo++) {
        if (HLOC(c, 0, 0, io, jo) != 0.0 && nup[jo] == 1 && nup[io] == 0) {
          int32_t k1, k2; double sg1, sg2;
          orc_c(jo + 1, mup, &k1, &sg1);
          orc_cdg(io + 1, k1, &k2, &sg2);
          int64_t j = (orc_binary_search(hu, du, k2) - 1) + idw * du;
          o->eknot += HLOC(c, 0, 0, io, jo) * sg1 * sg2 * gs[i] * gs[j] * peso;
        }
        if (HLOC(c, sl, sl, io, jo) != 0.0 && ndw[jo] == 1 && ndw[io] == 0) {
          int32_t k1, k2; double sg1, sg2;
          orc_c(jo + 1, mdw, &k1, &sg1);
          orc_cdg(io + 1, k1, &k2, &sg2);
          int64_t j = iup + (orc_binary_search(hd, dd, k2) - 1) * du;
          o->eknot += HLOC(c, sl, sl, io, jo) * sg1 * sg2 * gs[i] * gs[j] * peso;
        }
      }
    if (c->jhflag && c->jx != 0.0)
      for (int io = 0; io < norb; io++)
        for (int jo = 0; jo < norb; jo++)
          if (io != jo && nup[jo] == 1 && ndw[io] == 1 && ndw[jo] == 0 && nup[io] == 0) {
            int32_t k1, k2, k3, k4; double sg1, sg2, sg3, sg4;
            orc_c(io + 1, mdw, &k1, &sg1); orc_cdg(jo + 1, k1, &k2, &sg2);
            orc_c(jo + 1, mup, &k3, &sg3); orc_cdg(io + 1, k3, &k4, &sg4);
            int64_t j = (orc_binary_search(hu, du, k4) - 1) + (orc_binary_search(hd, dd, k2) - 1) * du;
            o->epot += c->jx * sg1 * sg2 * sg3 * sg4 * gs[i] * gs[j] * peso;
            o->dse += sg1 * sg2 * sg3 * sg4 * gs[i] * gs[j] * peso;
          }
    if (c->jhflag && c->jp != 0.0)
      for (int io = 0; io < norb; io++)
        for (int jo = 0; jo < norb; jo++)
          if (nup[jo] == 1 && ndw[jo] == 1 && ndw[io] == 0 && nup[io] == 0) {
            int32_t k1, k2, k3, k4; double sg1, sg2, sg3, sg4;
            orc_c(jo + 1, mdw, &k1, &sg1); orc_cdg(io + 1, k1, &k2, &sg2);
            orc_c(jo + 1, mup, &k3, &sg3); orc_cdg(io + 1, k3, &k4, &sg4);
            int64_t j = (orc_binary_search(hu, du, k4) - 1) + (orc_binary_search(hd, dd, k2) - 1) * du;
            o->epot += c->jp * sg1 * sg2 * sg3 * sg4 * gs[i] * gs[j] * peso;
            o->dph += sg1 * sg2 * sg3 * sg4 * gs[i] * gs[j] * peso;
          }
    for (int io = 0; io < norb; io++) o->epot += c->uloc[io] * nup[io] * ndw[io] * gs_weight;
    if (norb > 1) {
      for (int io = 0; io < norb; io++)
        for (int jo = io + 1; jo < norb; jo++) {
          o->epot += c->ust * (nup[io] * ndw[jo] + nup[jo] * ndw[io]) * gs_weight;
          o->dust += (nup[io] * ndw[jo] + nup[jo] * ndw[io]) * gs_weight;
        }
      for (int io = 0; io < norb; io++)
        for (int jo = io + 1; jo < norb; jo++) {
          o->epot += (c->ust - c->jh) * (nup[io] * nup[jo] + ndw[io] * ndw[jo]) * gs_weight;
          o->dund += (nup[io] * nup[jo] + ndw[io] * ndw[jo]) * gs_weight;
        }
    }
    if (c->hfmode) {
      for (int io = 0; io < norb; io++)
        o->ehartree += -0.5 * c->uloc[io] * (nup[io] + ndw[io]) * gs_weight + 0.25 * c->uloc[io] * gs_weight;
      if (norb > 1)
        for (int io = 0; io < norb; io++)
          for (int jo = io + 1; jo < norb; jo++) {
            o->ehartree += -0.5 * c->ust * (nup[io] + ndw[io] + nup[jo] + ndw[jo]) * gs_weight + 0.25 * c->ust * gs_weight;
            o->ehartree += -0.5 * (c->ust - c->jh) * (nup[io] + ndw[io] + nup[jo] + ndw[jo]) * gs_weight + 0.25 * (c->ust - c->jh) * gs_weight;
          }
    }
  }
  o->epot = o->epot + o->ehartree;                           /* :587 */
  free(hu); free(hd);
}

/* ===================================================================================== */
/* ED_GF_NORMAL                                                                          */
/* ===================================================================================== */
int64_t orc_gf_start_vector(const orc_ctx *c, int nup, int ndw, const double *gs,
                            int iorb, int ispin, int add, double *vvinit, double *norm2,
                            int *jnup_out, int *jndw_out) {
  /* ED_GF_NORMAL.f90:173-216 (add) / 248-290 (remove); ed_total_ud=T so ialfa=1, iorb1=iorb */
  int jnup = nup, jndw = ndw;
  if (ispin == 1) jnup += add ? 1 : -1; else jndw += add ? 1 : -1;
  if (jnup_out) *jnup_out = jnup;
  if (jndw_out) *jndw_out = jndw;
  if (jnup < 0 || jnup > c->ns || jndw < 0 || jndw > c->ns) return 0;  /* getC(DG)sector = 0 */
  int64_t idu = orc_binomial(c->ns, nup), idd = orc_binomial(c->ns, ndw);
  int64_t jdu = orc_binomial(c->ns, jnup), jdd = orc_binomial(c->ns, jndw);
  int64_t idim = idu * idd, jdim = jdu * jdd;
  if (!vvinit) return jdim;
  int32_t *hi_up = (int32_t *)xmalloc((size_t)idu * 4), *hi_dw = (int32_t *)xmalloc((size_t)idd * 4);
  int32_t *hj_up = (int32_t *)xmalloc((size_t)jdu * 4), *hj_dw = (int32_t *)xmalloc((size_t)jdd * 4);
  orc_build_sector_map(c->ns, nup, hi_up); orc_build_sector_map(c->ns, ndw, hi_dw);
  orc_build_sector_map(c->ns, jnup, hj_up); orc_build_sector_map(c->ns, jndw, hj_dw);
  for (int64_t j = 0; j < jdim; j++) vvinit[j] = 0.0;
  for (int64_t i = 0; i < idim; i++) {
    int64_t iu = i % idu, id = i / idu;
    int32_t m = (ispin == 1) ? hi_up[iu] : hi_dw[id];
    int occ = (m >> (iorb - 1)) & 1;
    int32_t r; double sgn;
    if (add) { if (occ != 0) continue; orc_cdg(iorb, m, &r, &sgn); }
    else     { if (occ != 1) continue; orc_c(iorb, m, &r, &sgn); }
    int64_t ju = iu, jd = id;
    if (ispin == 1) ju = orc_binary_search(hj_up, jdu, r) - 1;
    else            jd = orc_binary_search(hj_dw, jdd, r) - 1;
    vvinit[ju + jd * jdu] = sgn * gs[i];
  }
  double n2 = ddot(jdim, vvinit, vvinit);
  double sq = sqrt(n2);
  for (int64_t j = 0; j < jdim; j++) vvinit[j] = vvinit[j] / sq;
  *norm2 = n2;
  free(hi_up); free(hi_dw); free(hj_up); free(hj_dw);
  return jdim;
}

void orc_add_to_lanczos_gf(double norm2, double zeta, double ei, const double *alanc,
                           const double *blanc, int nlanc, int isign,
                           const double *wm, int lmats, double *gmats,
                           const double *wr, int lreal, double eps, double *greal) {
  /* ED_GF_NORMAL.f90:599-654, T=0: pesoBZ = vnorm2/zeta_function */
  double pesobz = norm2 / zeta;
  double *diag = (double *)xmalloc((size_t)nlanc * sizeof(double));
  double *z = (double *)xmalloc((size_t)nlanc * nlanc * sizeof(double));
  tridiag_eig(nlanc, alanc, blanc, diag, z);
  double complex *gm = (double complex *)gmats, *gr = (double complex *)greal;
  for (int j = 0; j < nlanc; j++) {
    double de = diag[j] - ei;
    double z1 = z[0 + (size_t)nlanc * j];
    double peso = pesobz * z1 * z1;
    for (int i = 0; i < lmats; i++) {
      double complex iw = I * wm[i];
      gm[i] = gm[i] + peso / (iw - (double)isign * de);
    }
    for (int i = 0; i < lreal; i++) {
      double complex iw = wr[i] + I * eps;
      gr[i] = gr[i] + peso / (iw - (double)isign * de);
    }
  }
  free(diag); free(z);
}

void orc_lanc_build_gf_normal_main(const orc_ctx *c, int nup, int ndw, const double *gs,
                                   double e0, double zeta, int iorb, int ispin, int ngfiter,
                                   int mode, const double *wm, int lmats, double *gmats,
                                   const double *wr, int lreal, double eps, double *greal,
                                   double *chain_out, int *nlanc_out) {
  for (int pass = 0; pass < 2; pass++) {                  /* 0: add (:173-246), 1: remove (:248-318) */
    int add = (pass == 0);
    int jnup, jndw; double norm2 = 0.0;
    int64_t jdim = orc_gf_start_vector(c, nup, ndw, gs, iorb, ispin, add, NULL, &norm2, &jnup, &jndw);
    if (nlanc_out) nlanc_out[pass] = 0;
    if (jdim == 0) continue;
    double *vv = (double *)xmalloc((size_t)jdim * sizeof(double));
    orc_gf_start_vector(c, nup, ndw, gs, iorb, ispin, add, vv, &norm2, &jnup, &jndw);
    int nlanc = (jdim < ngfiter) ? (int)jdim : ngfiter;
    double *alfa = (double *)xcalloc((size_t)nlanc, sizeof(double));
    double *beta = (double *)xcalloc((size_t)nlanc, sizeof(double));
    orc_sector *s = orc_build_hv_sector(c, jnup, jndw, 0, 1, mode == 0);
    orc_lanc_tridiag_sector(s, mode, vv, alfa, beta, nlanc, 1.0e-12);
    orc_delete_hv_sector(s);
    orc_add_to_lanczos_gf(norm2, zeta, e0, alfa, beta, nlanc, add ? 1 : -1,
                          wm, lmats, gmats, wr, lreal, eps, greal);
    if (chain_out) {
      double *co = chain_out + (size_t)pass * (1 + 2 * (size_t)ngfiter);
      co[0] = norm2;
      memcpy(co + 1, alfa, (size_t)nlanc * sizeof(double));
      memcpy(co + 1 + ngfiter, beta, (size_t)nlanc * sizeof(double));
    }
    if (nlanc_out) nlanc_out[pass] = nlanc;
    free(vv); free(alfa); free(beta);
  }
}

/* ===================================================================================== */
/* DimPh > 1: one local phonon mode (stored/H_ph.f90, H_e_ph.f90; spMatVec_main :391-485)  */
/* ===================================================================================== */
/* v(i_el, iph), i = i_el + (iph-1)*DimUp*DimDw, iph = 1..DimPh = Nph+1.  Term order of spMatVec_main: diagonal for every
 * element; then per phonon slab: dw hops, up hops, phonon energy w0*(iph-1), electron-phonon coupling
 * [sum_orb g_orb (n_up + n_dw - 1)] * (b + b^+) with the entries of spH0ph_eph in insertion order (destruction first:
 * column iph+1 with sqrt(iph), then construction: column iph-1 with sqrt(iph-1)); non-local terms last.  Serial. */
void orc_spmatvec_main_ph(const orc_sector *s, int nph, const double *g_ph, double w0_ph, int64_t nloc, const double *v, double *hv) {
  const int64_t du = s->dimup, dd = s->dimdw, nel = du * dd;
  const int dimph = nph + 1;
  const orc_ctx *c = s->ctx;
  if (nloc != nel * dimph) { fprintf(stderr, "spMatVec_main: Nloc != dim*DimPh\n"); abort(); }
  /* spH0e_eph, stored/H_e_ph.f90:1-26 */
  double *eeph = (double *)xmalloc((size_t)nel * sizeof(double));
  for (int64_t i = 0; i < nel; i++) {
    int nu[64], nd[64];
    bdecomp(s->map_up[i % du], c->ns, nu);
    bdecomp(s->map_dw[i / du], c->ns, nd);
    double htmp = 0.0;
    for (int io = 0; io < c->norb; io++) htmp = htmp + g_ph[io] * ((double)(nu[io] + nd[io]) - 1.0);
    eeph[i] = htmp;
  }
  for (int64_t i = 0; i < nloc; i++) hv[i] = 0.0;
  for (int64_t i = 0; i < nloc; i++) hv[i] = hv[i] + s->h0d[i % nel] * v[i];
  for (int iph = 0; iph < dimph; iph++) {
    const int64_t off = (int64_t)iph * nel;
    for (int64_t iup = 0; iup < du; iup++)
      for (int64_t idw = 0; idw < dd; idw++) {
        int64_t i = iup + idw * du + off;
        for (int64_t jj = s->hdw.rowptr[idw]; jj < s->hdw.rowptr[idw + 1]; jj++)
          hv[i] = hv[i] + s->hdw.vals[jj] * v[iup + s->hdw.cols[jj] * du + off];
      }
    for (int64_t idw = 0; idw < dd; idw++)
      for (int64_t iup = 0; iup < du; iup++) {
        int64_t i = iup + idw * du + off;
        for (int64_t jj = s->hup.rowptr[iup]; jj < s->hup.rowptr[iup + 1]; jj++)
          hv[i] = hv[i] + s->hup.vals[jj] * v[s->hup.cols[jj] + idw * du + off];
      }
    if (dimph > 1)
      for (int64_t ie = 0; ie < nel; ie++) {
        int64_t i = ie + off;
        hv[i] = hv[i] + (w0_ph * (double)iph) * v[i];                                   /* spH0_ph, H_ph.f90 */
        if (iph + 1 < dimph) hv[i] = hv[i] + (eeph[ie] * sqrt((double)(iph + 1))) * v[ie + off + nel];   /* b   */
        if (iph > 0) hv[i] = hv[i] + (eeph[ie] * sqrt((double)iph)) * v[ie + off - nel];                 /* b^+ */
      }
  }
  if (c->jhflag)
    for (int64_t i = 0; i < nloc; i++)
      for (int64_t jj = s->hnd.rowptr[i % nel]; jj < s->hnd.rowptr[i % nel + 1]; jj++)
        hv[i] = hv[i] + s->hnd.vals[jj] * v[s->hnd.cols[jj] + (i / nel) * nel];
  free(eeph);
}
typedef struct { const orc_sector *s; int nph; const double *g; double w0; } ph_mv;
static void ph_matvec(void *u, int64_t n, const double *v, double *hv) {
  ph_mv *m = (ph_mv *)u;
  orc_spmatvec_main_ph(m->s, m->nph, m->g, m->w0, n, v, hv);
}
int orc_lanc_eigh_sector_ph(const orc_sector *s, int nph, const double *g_ph, double w0_ph, double *egs, double *vect,
                            int nitermax, double threshold, int ncheck, int *nlanc_out, double *alanc_out, double *blanc_out) {
  ph_mv m = { s, nph, g_ph, w0_ph };
  return orc_sp_lanc_eigh(ph_matvec, &m, s->dim * (nph + 1), egs, vect, nitermax, threshold, ncheck, nlanc_out, alanc_out, blanc_out);
}

/* ===================================================================================== */
/* ed_total_ud = F: one (Nup, Ndw) pair per orbital (Ns_Ud = Norb, Ns_Orb = 1 + Nbath)    */
/* ===================================================================================== */
/* get_Sector, ED_SETUP.f90:446-457 with QN = [Nups, Ndws] and N = Ns_Orb */
int orc_get_sector_orbs(const orc_ctx *c, const int *nups, const int *ndws) {
  const int nind = 2 * c->norb, factor = c->nbath + 1 + 1;
  int isector = 1;
  for (int i = nind; i >= 1; i--) {
    int qn = (i <= c->norb) ? nups[i - 1] : ndws[i - 1 - c->norb];
    int pw = 1;
    for (int k = 0; k < nind - i; k++) pw *= factor;
    isector = isector + qn * pw;
  }
  return isector;
}

/* ed_buildh_orbs (ED_HAMILTONIAN_SPARSE_HxV.f90:206-370) with stored/Orbs/H_local.f90, H_up.f90, H_dw.f90: per
 * orbital and spin the star of that orbital's impurity site (bit 0 of the orbital word) and its own bath levels (bits
 * 1..Nbath); the diagonal from the occupations reordered to the Ns site numbering (breorder, ED_SETUP.f90:963-979:
 * impurity orbitals first, then the bath levels of orbital 1, 2, ...).  Serial only (rank 0 of 1). */
orc_sector_orbs *orc_build_hv_sector_orbs(const orc_ctx *c, const int *nups, const int *ndws) {
  orc_sector_orbs *s = (orc_sector_orbs *)xcalloc(1, sizeof(*s));
  const int norb = c->norb, nso = c->nbath + 1;
  s->ctx = c; s->nfac = 2 * norb; s->dim = 1;
  for (int f = 0; f < 2 * norb; f++) {
    int n = (f < norb) ? nups[f] : ndws[f - norb];
    s->nq[f] = n;
    s->dims[f] = orc_binomial(nso, n);
    s->map[f] = (int32_t *)xmalloc((size_t)s->dims[f] * 4);
    orc_build_sector_map(nso, n, s->map[f]);
    s->dim *= s->dims[f];
  }
  /* H_up / H_dw of every orbital, insertion order: outer loop over the source state, kp inner (Orbs/H_up.f90) */
  for (int f = 0; f < 2 * norb; f++) {
    const int io = f % norb, is = (f < norb) ? 0 : c->nspin - 1;
    rl_mat m;
    rl_init(&m, s->dims[f]);
    for (int64_t j = 0; j < s->dims[f]; j++) {
      int32_t mm = s->map[f][j];
      int occ[64];
      bdecomp(mm, nso, occ);
      for (int kp = 1; kp <= c->nbath; kp++) {
        int ialfa = 1 + kp;
        double v = BATHV(c, is, io, kp - 1);
        if (v != 0.0 && occ[0] == 1 && occ[ialfa - 1] == 0) {
          int32_t k1, k2; double sg1, sg2;
          orc_c(1, mm, &k1, &sg1); orc_cdg(ialfa, k1, &k2, &sg2);
          int64_t i = orc_binary_search(s->map[f], s->dims[f], k2) - 1;
          rl_insert(&m, v * sg1 * sg2, i, j);
        }
        if (v != 0.0 && occ[0] == 0 && occ[ialfa - 1] == 1) {
          int32_t k1, k2; double sg1, sg2;
          orc_c(ialfa, mm, &k1, &sg1); orc_cdg(1, k1, &k2, &sg2);
          int64_t i = orc_binary_search(s->map[f], s->dims[f], k2) - 1;
          rl_insert(&m, v * sg1 * sg2, i, j);
        }
      }
    }
    rl_to_csr(&m, &s->h[f]);
  }
  /* diagonal */
  s->h0d = (double *)xmalloc((size_t)s->dim * sizeof(double));
  for (int64_t i = 0; i < s->dim; i++) {
    int nup[64], ndw[64];
    int64_t count = i;
    for (int k = 0; k < c->ns; k++) { nup[k] = 0; ndw[k] = 0; }
    for (int f = 0; f < 2 * norb; f++) {
      int64_t idx = count % s->dims[f];
      count /= s->dims[f];
      int occ[64];
      bdecomp(s->map[f][idx], nso, occ);
      int io = f % norb;
      int *dst = (f < norb) ? nup : ndw;
      dst[io] = occ[0];
      for (int kp = 1; kp <= c->nbath; kp++) dst[orc_bath_stride(c, io + 1, kp) - 1] = occ[kp];
    }
    s->h0d[i] = h_local_element(c, nup, ndw);
  }
  return s;
}

void orc_delete_hv_sector_orbs(orc_sector_orbs *s) {
  if (!s) return;
  for (int f = 0; f < s->nfac; f++) { free(s->map[f]); csr_free(&s->h[f]); }
  free(s->h0d);
  free(s);
}

/* spMatVec_orbs, ED_HAMILTONIAN_SPARSE_HxV.f90:487-564 (DimPh = 1): diagonal, then per element and per orbital the up
 * entries followed by the dw entries. */
void orc_spmatvec_orbs(const orc_sector_orbs *s, int64_t nloc, const double *v, double *hv) {
  const int norb = s->nfac / 2;
  for (int64_t i = 0; i < nloc; i++) hv[i] = 0.0;
  for (int64_t i = 0; i < nloc; i++) hv[i] = hv[i] + s->h0d[i] * v[i];
  int64_t stride[2 * ORC_MAX_ORB];
  stride[0] = 1;
  for (int f = 1; f < s->nfac; f++) stride[f] = stride[f - 1] * s->dims[f - 1];
  for (int64_t i = 0; i < nloc; i++) {
    int64_t idx[2 * ORC_MAX_ORB], count = i;
    for (int f = 0; f < s->nfac; f++) { idx[f] = count % s->dims[f]; count /= s->dims[f]; }
    for (int iud = 0; iud < norb; iud++) {
      for (int sp = 0; sp < 2; sp++) {                       /* UP then DW of this orbital */
        const int f = iud + sp * norb;
        const orc_csr *h = &s->h[f];
        for (int64_t p = h->rowptr[idx[f]]; p < h->rowptr[idx[f] + 1]; p++) {
          int64_t j = i + (h->cols[p] - idx[f]) * stride[f];
          hv[i] = hv[i] + h->vals[p] * v[j];
        }
      }
    }
  }
}

static void orbs_matvec(void *u, int64_t n, const double *v, double *hv) { orc_spmatvec_orbs((const orc_sector_orbs *)u, n, v, hv); }
int orc_lanc_eigh_sector_orbs(const orc_sector_orbs *s, double *egs, double *vect, int nitermax, double threshold, int ncheck,
                              int *nlanc_out, double *alanc_out, double *blanc_out) {
  return orc_sp_lanc_eigh(orbs_matvec, (void *)s, s->dim, egs, vect, nitermax, threshold, ncheck, nlanc_out, alanc_out, blanc_out);
}
int orc_lanc_tridiag_sector_orbs(const orc_sector_orbs *s, double *vin, double *alanc, double *blanc, int nitermax, double threshold) {
  return orc_sp_lanc_tridiag(orbs_matvec, (void *)s, s->dim, vin, alanc, blanc, nitermax, threshold);
}
void orc_sector_orbs_info(const orc_sector_orbs *s, int64_t *dim, int64_t *dims) {
  *dim = s->dim;
  for (int f = 0; f < s->nfac; f++) dims[f] = s->dims[f];
}
void orc_sector_orbs_get(const orc_sector_orbs *s, int f, int32_t *map, int64_t *rowptr, int64_t *cols, double *vals, double *h0d) {
  if (map) memcpy(map, s->map[f], (size_t)s->dims[f] * 4);
  if (rowptr) memcpy(rowptr, s->h[f].rowptr, (size_t)(s->dims[f] + 1) * 8);
  if (cols) memcpy(cols, s->h[f].cols, (size_t)s->h[f].rowptr[s->dims[f]] * 8);
  if (vals) memcpy(vals, s->h[f].vals, (size_t)s->h[f].rowptr[s->dims[f]] * 8);
  if (h0d) memcpy(h0d, s->h0d, (size_t)s->dim * 8);
}

/* ===================================================================================== */
/* ED_GF_CHISPIN / ED_GF_CHIDENS (ed_total_ud = T: ialfa = 1, iorb1 = iorb)              */
/* ===================================================================================== */
/* start vector O|gs> in the state's own sector, then norm2 and normalisation.
 * kind 0 (spin): lanc_ed_build_spinChi_main ED_GF_CHISPIN.f90:150-170 (iorb == jorb >= 1), _tot_main :255-287 (iorb == 0),
 *                _mix_main :370-400 (iorb != jorb);
 * kind 1 (dens): lanc_ed_build_densChi_main ED_GF_CHIDENS.f90:150-175, _tot_main :255-285, _mix_main :370-395. */
int64_t orc_chi_start_vector(const orc_ctx *c, int nup, int ndw, const double *gs, int kind, int iorb, int jorb,
                             double *vvinit, double *norm2) {
  int64_t idu = orc_binomial(c->ns, nup), idd = orc_binomial(c->ns, ndw), idim = idu * idd;
  int32_t *hi_up = (int32_t *)xmalloc((size_t)idu * 4), *hi_dw = (int32_t *)xmalloc((size_t)idd * 4);
  orc_build_sector_map(c->ns, nup, hi_up); orc_build_sector_map(c->ns, ndw, hi_dw);
  int nu[64], nd[64];
  for (int64_t i = 0; i < idim; i++) {
    int64_t iu = i % idu, id = i / idu;
    bdecomp(hi_up[iu], c->ns, nu);
    bdecomp(hi_dw[id], c->ns, nd);
    double sgn;
    if (iorb == 0) {                                           /* total: sum over the impurity orbitals */
      double sup = 0.0, sdw = 0.0;
      for (int a = 0; a < c->norb; a++) { sup += nu[a]; sdw += nd[a]; }
      sgn = kind == 0 ? 0.5 * (sup - sdw) : (sup + sdw);
    } else if (iorb == jorb) {
      sgn = kind == 0 ? 0.5 * ((double)nu[iorb - 1] - (double)nd[iorb - 1]) : (double)(nu[iorb - 1] + nd[iorb - 1]);
    } else {
      double si = kind == 0 ? (double)(nu[iorb - 1] - nd[iorb - 1]) : (double)(nu[iorb - 1] + nd[iorb - 1]);
      double sj = kind == 0 ? (double)(nu[jorb - 1] - nd[jorb - 1]) : (double)(nu[jorb - 1] + nd[jorb - 1]);
      sgn = kind == 0 ? 0.5 * si + 0.5 * sj : si + sj;
    }
    vvinit[i] = sgn * gs[i];
  }
  double n2 = ddot(idim, vvinit, vvinit);
  double sq = sqrt(n2);
  for (int64_t i = 0; i < idim; i++) vvinit[i] = vvinit[i] / sq;
  *norm2 = n2;
  free(hi_up); free(hi_dw);
  return idim;
}

/* add_to_lanczos_spinChi ED_GF_CHISPIN.f90:434-488 == add_to_lanczos_densChi ED_GF_CHIDENS.f90:436-489, T = 0
 * (pesoBZ = 1).  chi_iv[0..lmats], chi_tau[0..ltau], chi_w[lreal] complex; accumulates. */
void orc_add_to_lanczos_chi(double norm2, double zeta, double ei, double beta, const double *alanc, const double *blanc, int nlanc,
                            const double *vm, int lmats, double *chi_iv, const double *tau, int ltau, double *chi_tau,
                            const double *vr, int lreal, double eps, double *chi_w) {
  double pesof = norm2 / zeta, pesobz = 1.0;
  double *diag = (double *)xmalloc((size_t)nlanc * sizeof(double));
  double *z = (double *)xmalloc((size_t)nlanc * nlanc * sizeof(double));
  tridiag_eig(nlanc, alanc, blanc, diag, z);
  double complex *cw = (double complex *)chi_w;
  for (int j = 0; j < nlanc; j++) {
    double de = diag[j] - ei;
    double z1 = z[0 + (size_t)nlanc * j];
    double peso = pesof * (z1 * z1) * pesobz;
    if (beta * de > 1e-3) chi_iv[0] += peso * 2 * (1.0 - exp(-beta * de)) / de;
    for (int i = 1; i <= lmats; i++) chi_iv[i] += peso * (1.0 - exp(-beta * de)) * 2.0 * de / (vm[i] * vm[i] + de * de);
    for (int i = 0; i <= ltau; i++) chi_tau[i] += exp(-tau[i] * de) * peso;
    for (int i = 0; i < lreal; i++) {
      double complex w = vr[i] + eps * I;
      cw[i] -= peso * (1.0 - exp(-beta * de)) * (1.0 / (w - de) - 1.0 / (w + de));
    }
  }
  free(diag); free(z);
}

/* one chain: start vector, sp_lanc_tridiag in the state's sector (nlanc = min(idim, lanc_nGFiter)) */
int orc_chi_chain(const orc_ctx *c, int nup, int ndw, const double *gs, int kind, int iorb, int jorb, int ngfiter, int mode,
                  double *norm2, double *alanc, double *blanc) {
  int64_t idim = (int64_t)orc_binomial(c->ns, nup) * orc_binomial(c->ns, ndw);
  double *vv = (double *)xmalloc((size_t)idim * sizeof(double));
  orc_chi_start_vector(c, nup, ndw, gs, kind, iorb, jorb, vv, norm2);
  int nlanc = (idim < ngfiter) ? (int)idim : ngfiter;
  for (int k = 0; k < nlanc; k++) { alanc[k] = 0.0; blanc[k] = 0.0; }
  orc_sector *s = orc_build_hv_sector(c, nup, ndw, 0, 1, mode == 0);
  orc_lanc_tridiag_sector(s, mode, vv, alanc, blanc, nlanc, 1.0e-12);
  orc_delete_hv_sector(s);
  free(vv);
  return nlanc;
}

/* Sigma = G0^-1 - G^-1, G0^-1 = z + xmu - impHloc - sum_k V_k^2/(z - e_k) */
void orc_sigma_normal(const orc_ctx *c, int iorb, int ispin, const double *zin, int l,
                      const double *g, double *sigma, double *invg0) {
  const double complex *z = (const double complex *)zin, *gg = (const double complex *)g;
  double complex *sg = (double complex *)sigma, *ig0 = (double complex *)invg0;
  int is = ispin - 1, io = iorb - 1;
  for (int i = 0; i < l; i++) {
    double complex delta = 0.0;
    for (int k = 0; k < c->nbath; k++) {
      double vps = BATHV(c, is, io, k), eps = BATHE(c, is, io, k);
      delta += vps * vps / (z[i] - eps);
    }
    double complex g0inv = z[i] + c->xmu - HLOC(c, is, is, io, io) - delta;
    if (ig0) ig0[i] = g0inv;
    sg[i] = g0inv - 1.0 / gg[i];
  }
}

/* allocate_grids, ED_AUX_FUNX.f90:281-295 (wm, wr only) */
void orc_allocate_grids(double beta, int lmats, double wini, double wfin, int lreal,
                        double *wm, double *wr) {
  const double pi = 3.14159265358979323846;
  for (int i = 1; i <= lmats; i++) wm[i - 1] = pi / beta * (double)(2 * i - 1);
  for (int i = 0; i < lreal; i++)
    wr[i] = (lreal > 1) ? wini + (wfin - wini) * (double)i / (double)(lreal - 1) : wini;
}
