/*
 * ed_oracle.h -- CPU ORACLE for the Lanczos H*v path of lcrippa/dmft-lanc-ed.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.  The product
 * path (dmft-lanc-ed_b200/csrc, libedgpu.so) never links, loads or calls it.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures and cannot be
 * compiled in this environment (no Fortran compiler, SciFortran not vendored).  This file
 * is a loop-order-faithful C restatement of the reference algorithm; it is pinned only by
 * the known-answer tests in tests/test_oracle_*.py (U=0 analytic, atomic limit, dense
 * Kronecker + LAPACK, stored == direct == sharded, hermiticity, GF sum rules).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).  Indices are 0-based here; the reference is 1-based.  Global state
 * indices are int64 (the reference's default INTEGER overflows at Ns=18, 9:9).
 */
#ifndef ED_ORACLE_H
#define ED_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_ORB 5

/* Module-global inputs read by build_Hv_sector (ED_INPUT_VARS.f90:129-208,
 * ED_VARS_GLOBAL.f90:105-146).  bath_type = "normal" only. */
typedef struct orc_ctx {
  int norb, nbath, nspin, ns;   /* ns = (nbath+1)*norb, ED_SETUP.f90:113-116 */
  int hfmode;
  int jhflag;                   /* ED_SETUP.f90:147-148 */
  double uloc[ORC_MAX_ORB], ust, jh, jx, jp, xmu;
  double *imphloc;              /* (nspin,nspin,norb,norb) Fortran order */
  double *bath_e;               /* dmft_bath%e(nspin,norb,nbath) Fortran order */
  double *bath_v;               /* dmft_bath%v(nspin,norb,nbath) Fortran order */
  /* bath_type (ED_SETUP.f90:113-121, 358-375): 0 normal, 1 hybrid (Ns = Norb + Nbath, every level shared by all
   * orbitals: bath_e only uses orbital column 1, nfoo = 1), 2 replica (Ns = Norb*(Nbath+1), level kp = one replica of
   * the impurity with its own Norb x Norb matrix bath_h(:,:,:,:,kp) and ONE hybridisation v(ispin,kp) for all orbitals;
   * bath_e then holds its diagonal and bath_v the replicated v) */
  int bath_type, nfoo;
  double *bath_h;               /* replica: Hbath(nspin,nspin,norb,norb,nbath) Fortran order, else NULL */
} orc_ctx;

/* Row-list sparse matrix flattened to CSR in insertion order (ED_SPARSE_MATRIX.f90:13-30). */
typedef struct orc_csr {
  int64_t nrow, ncol;
  int64_t *rowptr;              /* nrow+1 */
  int64_t *cols;                /* 0-based */
  double  *vals;
} orc_csr;

/* State created by build_Hv_sector (ED_HAMILTONIAN.f90:43-168, ED_HAMILTONIAN_COMMON.f90:11-18). */
typedef struct orc_sector {
  const orc_ctx *ctx;
  int nup, ndw;
  int64_t dimup, dimdw, dim;
  int32_t *map_up, *map_dw;     /* Hs(1)%map, Hs(2)%map */
  /* dw split, ED_HAMILTONIAN.f90:96-110 */
  int rank, nranks;
  int64_t qdw, rdw, q, r, istart, iend, ishift;   /* istart/iend 0-based half-open */
  int sparse_h;
  double *h0d;                  /* spH0d as one value per local row (stored/H_local.f90:73-78) */
  orc_csr hup, hdw, hnd;        /* spH0ups(1), spH0dws(1), spH0nd */
} orc_sector;

/* ---- setup ------------------------------------------------------------------------- */
orc_ctx *orc_ctx_create(int norb, int nbath, int nspin, int hfmode, const double *uloc,
                        double ust, double jh, double jx, double jp, double xmu,
                        const double *imphloc, const double *bath_e, const double *bath_v);
/* the same for bath_type 1 (hybrid: bath_e(nspin,1,nbath), bath_v(nspin,norb,nbath)) and 2 (replica: bath_e unused,
 * bath_v(nspin,nbath), bath_h(nspin,nspin,norb,norb,nbath) = bath_from_sym(lambda) as ed_buildh_main assembles it,
 * ED_HAMILTONIAN_SPARSE_HxV.f90:61-75) */
orc_ctx *orc_ctx_create_bt(int norb, int nbath, int nspin, int hfmode, const double *uloc,
                           double ust, double jh, double jx, double jp, double xmu, const double *imphloc,
                           int bath_type, const double *bath_e, const double *bath_v, const double *bath_h);
void orc_ctx_destroy(orc_ctx *c);
void orc_init_dmft_bath(int norb, int nbath, int nspin, double hwband, double *e, double *v);

/* ---- ED_SETUP hot subset ------------------------------------------------------------ */
int     orc_binomial(int n1, int n2);
int64_t orc_build_sector_map(int ns, int n, int32_t *map);   /* map may be NULL: count only */
int     orc_c(int pos, int32_t in, int32_t *out, double *fsgn);
int     orc_cdg(int pos, int32_t in, int32_t *out, double *fsgn);
int64_t orc_binary_search(const int32_t *a, int64_t n, int32_t value); /* 1-based, 0 = absent */
int     orc_get_sector(int nup, int ndw, int ns);            /* 1-based isector */
void    orc_get_nup_ndw(int isector, int ns, int *nup, int *ndw);
int     orc_bath_stride(const orc_ctx *c, int iorb, int kp); /* 1-based site */

/* ---- ED_HAMILTONIAN ----------------------------------------------------------------- */
orc_sector *orc_build_hv_sector(const orc_ctx *c, int nup, int ndw, int rank, int nranks,
                                int sparse_h);
void    orc_delete_hv_sector(orc_sector *s);
int64_t orc_vecdim_hv_sector(const orc_ctx *c, int nup, int ndw, int rank, int nranks);
void    orc_build_hmat(const orc_sector *s, double *hmat);   /* dense, column-major dim x dim */

void orc_spmatvec_main(const orc_sector *s, int64_t nloc, const double *v, double *hv);
void orc_directmatvec_main(const orc_sector *s, int64_t nloc, const double *v, double *hv);
/* spMatVec_main loops restricted to the column block of sector view s (rank r of P), full v in */
void orc_spmatvec_block(const orc_sector *s, const double *v_full, double *hv_block);
int64_t orc_spmatvec_blocks_mt(orc_sector **secs, int nblk, int nthreads, const double *v_full,
                               double **hv_blocks);
/* same loops, for a caller that holds only the columns of v the block touches (ascending colidx) */
int orc_spmatvec_block_cols(const orc_sector *s, const int64_t *colidx, int64_t ncolidx, const double *xcols,
                            double *hv_block);
/* Emulation of nranks MPI ranks in one process; v/hv are the concatenated shards (= the
 * serial vector, because the split is by contiguous i_dw column blocks). */
void orc_spmatvec_mpi_main_all(const orc_ctx *c, int nup, int ndw, int nranks, int nthreads,
                               const double *v, double *hv);
void orc_directmatvec_mpi_main_all(const orc_ctx *c, int nup, int ndw, int nranks, int nthreads,
                                   const double *v, double *hv);
/* same, with pre-built per-rank sectors (for timing without the build) */
void orc_spmatvec_mpi_main_prebuilt(orc_sector **secs, int nranks, int nthreads,
                                    const double *v, double *hv);
/* ---- observables of one state at T = 0 (lanc_observables, ED_OBSERVABLES.f90:95-363, and lanc_local_energy,
 * :372-600; bath_type normal, ed_total_ud = T, DimPh = 1).  Arrays use the Fortran index order of the reference
 * with leading dimension ORC_MAX_ORB: sz2[iorb + 5*jorb], dm[ispin][iorb + 5*jorb] =
 * imp_density_matrix(ispin,ispin,iorb,jorb) for ispin <= Nspin. */
typedef struct orc_observables {
  double dens[ORC_MAX_ORB], dens_up[ORC_MAX_ORB], dens_dw[ORC_MAX_ORB], docc[ORC_MAX_ORB], magz[ORC_MAX_ORB];
  double sz2[ORC_MAX_ORB * ORC_MAX_ORB], n2[ORC_MAX_ORB * ORC_MAX_ORB];
  double s2tot;
  double prob[243];
  double dm[2][ORC_MAX_ORB * ORC_MAX_ORB];
  double eknot, epot, ehartree, dust, dund, dse, dph;
} orc_observables;
void orc_observables_normal(const orc_ctx *c, int nup, int ndw, const double *gs, double zeta, orc_observables *out);

/* vector_transpose_MPI restated for all ranks at once (ED_HAMILTONIAN_COMMON.f90:53-118). */
void orc_vector_transpose_all(int nranks, int64_t nrow, int64_t ncol,
                              double *const *a_shards, double *const *b_shards);

/* ---- SciFortran simple Lanczos (external, restated; see ed_oracle.c header note) ---- */
typedef void (*orc_matvec_fn)(void *user, int64_t nloc, const double *v, double *hv);
int orc_tql2(int n, double *d, double *e, double *z);        /* z column-major n x n, in: identity */
int orc_sp_lanc_eigh(orc_matvec_fn mv, void *user, int64_t n, double *egs, double *vect,
                     int nitermax, double threshold, int ncheck,
                     int *nlanc_out, double *alanc_out, double *blanc_out);
int orc_sp_lanc_tridiag(orc_matvec_fn mv, void *user, int64_t n, double *vin,
                        double *alanc, double *blanc, int nitermax, double threshold);
/* convenience: mode 0 = spMatVec_main, 1 = directMatVec_main */
int orc_lanc_eigh_sector(const orc_sector *s, int mode, double *egs, double *vect,
                         int nitermax, double threshold, int ncheck,
                         int *nlanc_out, double *alanc_out, double *blanc_out);
int orc_lanc_tridiag_sector(const orc_sector *s, int mode, double *vin,
                            double *alanc, double *blanc, int nitermax, double threshold);

/* ---- ED_GF_NORMAL ------------------------------------------------------------------- */
/* Start vector c^+_{iorb,ispin}|gs> (add=1) or c_{iorb,ispin}|gs> (add=0):
 * ED_GF_NORMAL.f90:184-216 / 259-290.  Returns target dim (0 if the sector does not exist);
 * vvinit (length jdim) is normalised, *norm2 = <v|v> before normalisation. */
int64_t orc_gf_start_vector(const orc_ctx *c, int nup, int ndw, const double *gs,
                            int iorb, int ispin, int add, double *vvinit, double *norm2,
                            int *jnup, int *jndw);
/* add_to_lanczos_gf_normal, ED_GF_NORMAL.f90:599-654 (T=0 branch).  g arrays are
 * interleaved (re,im). */
void orc_add_to_lanczos_gf(double norm2, double zeta, double ei, const double *alanc,
                           const double *blanc, int nlanc, int isign,
                           const double *wm, int lmats, double *gmats,
                           const double *wr, int lreal, double eps, double *greal);
/* Full chain for one (iorb,ispin): both add and remove; ED_GF_NORMAL.f90:124-334. */
void orc_lanc_build_gf_normal_main(const orc_ctx *c, int nup, int ndw, const double *gs,
                                   double e0, double zeta, int iorb, int ispin, int ngfiter,
                                   int mode, const double *wm, int lmats, double *gmats,
                                   const double *wr, int lreal, double eps, double *greal,
                                   double *chain_out /* [2][1+2*ngfiter]: norm2, a[], b[] */,
                                   int *nlanc_out /* [2] */);
/* build_sigma_normal for bath_type normal, one (ispin,iorb): ED_GF_NORMAL.f90:935-1002,
 * ED_BATH_FUNCTIONS.f90:43-77,163-195. z interleaved complex. */
void orc_sigma_normal(const orc_ctx *c, int iorb, int ispin, const double *z, int l,
                      const double *g, double *sigma, double *invg0);
void orc_allocate_grids(double beta, int lmats, double wini, double wfin, int lreal,
                        double *wm, double *wr);
/* ---- DimPh = Nph + 1 > 1: one local phonon mode (stored/H_ph.f90, H_e_ph.f90; spMatVec_main :391-485).  v(i_el, iph),
 * phonon index slowest.  s: the electron sector (serial, stored diagonal). */
void orc_spmatvec_main_ph(const orc_sector *s, int nph, const double *g_ph, double w0_ph, int64_t nloc, const double *v, double *hv);
int orc_lanc_eigh_sector_ph(const orc_sector *s, int nph, const double *g_ph, double w0_ph, double *egs, double *vect,
                            int nitermax, double threshold, int ncheck, int *nlanc_out, double *alanc_out, double *blanc_out);
/* ---- ed_total_ud = F (Ns_Ud = Norb): ed_buildh_orbs / spMatVec_orbs, ED_HAMILTONIAN_SPARSE_HxV.f90:206-370, 487-564,
 * with ED_HAMILTONIAN/stored/Orbs/H_local.f90, H_up.f90, H_dw.f90.  Factor f < Norb: up word of orbital f+1, else the
 * dw word of orbital f+1-Norb; the vector index runs over [DimUps, DimDws] with the first factor fastest
 * (state2indices, ED_SETUP.f90:520-545).  Serial. */
typedef struct orc_sector_orbs {
  const orc_ctx *ctx;
  int nfac, nq[2 * ORC_MAX_ORB];
  int64_t dims[2 * ORC_MAX_ORB], dim;
  int32_t *map[2 * ORC_MAX_ORB];
  orc_csr h[2 * ORC_MAX_ORB];   /* spH0ups(iud), spH0dws(iud) */
  double *h0d;                  /* spH0d, one value per state */
} orc_sector_orbs;
int orc_get_sector_orbs(const orc_ctx *c, const int *nups, const int *ndws);
orc_sector_orbs *orc_build_hv_sector_orbs(const orc_ctx *c, const int *nups, const int *ndws);
void orc_delete_hv_sector_orbs(orc_sector_orbs *s);
void orc_spmatvec_orbs(const orc_sector_orbs *s, int64_t nloc, const double *v, double *hv);
int orc_lanc_eigh_sector_orbs(const orc_sector_orbs *s, double *egs, double *vect, int nitermax, double threshold, int ncheck,
                              int *nlanc_out, double *alanc_out, double *blanc_out);
int orc_lanc_tridiag_sector_orbs(const orc_sector_orbs *s, double *vin, double *alanc, double *blanc, int nitermax, double threshold);
void orc_sector_orbs_info(const orc_sector_orbs *s, int64_t *dim, int64_t *dims);
void orc_sector_orbs_get(const orc_sector_orbs *s, int f, int32_t *map, int64_t *rowptr, int64_t *cols, double *vals, double *h0d);
/* Susceptibility chains, ED_GF_CHISPIN.f90:114-415 / ED_GF_CHIDENS.f90:111-426 (ed_total_ud = T): start vector
 * O|gs> in the state's own sector.  kind 0 = spin (O = Sz), 1 = density (O = n); iorb == jorb >= 1: one orbital,
 * iorb == 0: total over the impurity orbitals, iorb != jorb: the mixed combination O_i + O_j. */
int64_t orc_chi_start_vector(const orc_ctx *c, int nup, int ndw, const double *gs, int kind, int iorb, int jorb,
                             double *vvinit, double *norm2);
int orc_chi_chain(const orc_ctx *c, int nup, int ndw, const double *gs, int kind, int iorb, int jorb, int ngfiter, int mode,
                  double *norm2, double *alanc, double *blanc);
/* add_to_lanczos_spinChi / _densChi (identical), T = 0.  vm[0..lmats], tau[0..ltau], vr[lreal]; chi_w interleaved. */
void orc_add_to_lanczos_chi(double norm2, double zeta, double ei, double beta, const double *alanc, const double *blanc, int nlanc,
                            const double *vm, int lmats, double *chi_iv, const double *tau, int ltau, double *chi_tau,
                            const double *vr, int lreal, double eps, double *chi_w);

#ifdef __cplusplus
}
#endif
#endif
