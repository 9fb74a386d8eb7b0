/*
 * edgpu.h -- C-ABI of the B200-native Lanczos H*v engine for dmft-lanc-ed's N_up:N_dw solver.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Every entry
 * point names the reference interface it replaces (paths relative to the reference root).
 * All functions return 0 on success and a non-zero status on error (the Fortran shim turns a
 * non-zero status into `stop`, the reference's abort-on-error convention); the message is
 * available from edgpu_last_error().  Unless stated otherwise, arrays are host memory owned by
 * the caller; device buffers are owned by the handle.
 *
 * Layout conventions (ED_SETUP.f90:547-560): a sector vector is the column-major matrix
 * v(i_up, i_dw), i = i_up + i_dw*DimUp (0-based here).  With nranks>1 rank r owns the
 * contiguous block of Q_r = DimDw/P (+1 for r < mod(DimDw,P)) i_dw columns
 * (ED_HAMILTONIAN.f90:96-110); "nloc" below is always DimUp*Q_r = vecDim_Hv_sector.
 *
 * There is NO CPU fallback: every compute entry fails with EDGPU_ERR_NO_DEVICE when no
 * sm_100 device is usable.
 */
#ifndef EDGPU_H
#define EDGPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EDGPU_MAX_ORB 5

enum {
  EDGPU_OK = 0,
  EDGPU_ERR_INVALID = 1,      /* bad argument / lifecycle misuse (reference: stop "...") */
  EDGPU_ERR_NO_DEVICE = 2,    /* no usable CUDA device: the product never falls back to CPU */
  EDGPU_ERR_CUDA = 3,
  EDGPU_ERR_NCCL = 4,
  EDGPU_ERR_UNSUPPORTED = 5   /* DimPh>1, ed_total_ud=F, bath_type/=normal */
};

/* H*v algorithm selector (edgpu_set_option "hxv_algo") */
enum {
  EDGPU_ALGO_AUTO = 0,
  EDGPU_ALGO_GATHER = 1,      /* one pass, global-memory gathers (reference kernel of the engine) */
  EDGPU_ALGO_TILED = 2,       /* two passes, shared-memory staged column / row-tile kernels (generic factors) */
  EDGPU_ALGO_FAST = 3         /* two passes, TMA-staged whole-column kernel + structured row kernel */
};

/* The module-global inputs build_Hv_sector reads (ED_INPUT_VARS.f90:129-208,
 * ED_VARS_GLOBAL.f90:105-146, ED_HAMILTONIAN_SPARSE_HxV.f90:46-76).  Arrays are Fortran
 * order exactly as the reference holds them, so the shim can pass them unchanged. */
typedef struct edgpu_params {
  int32_t norb, nbath, nspin;          /* NORB, NBATH, NSPIN; Ns = (Nbath+1)*Norb */
  int32_t hfmode;                      /* HFMODE */
  int32_t ed_sparse_h;                 /* ED_SPARSE_H: 1 = stored (spMatVec_*), 0 = direct */
  int32_t nph;                         /* NPH: DimPh = NPH + 1 phonon states (0: no phonons) */
  int32_t ed_total_ud;                 /* ED_TOTAL_UD: 1 = total (Nup, Ndw) sectors, 0 = one pair per orbital */
  int32_t bath_type;                   /* BATH_TYPE: 0 normal, 1 hybrid (Ns = Norb + Nbath), 2 replica (ED_SETUP.f90:113-121, 358-375) */
  double  uloc[EDGPU_MAX_ORB];         /* ULOC */
  double  ust, jh, jx, jp, xmu;        /* UST, JH, JX, JP, XMU */
  const double *imphloc;               /* impHloc(Nspin,Nspin,Norb,Norb), may be NULL (=0) */
  const double *bath_e;                /* dmft_bath%e(Nspin,Norb,Nbath); hybrid: (Nspin,1,Nbath); replica: unused */
  const double *bath_v;                /* dmft_bath%v(Nspin,Norb,Nbath); replica: dmft_bath%item(k)%v(ispin) as (Nspin,Nbath) */
  double  g_ph[EDGPU_MAX_ORB];         /* G_PH: electron-phonon couplings (NPH > 0) */
  double  w0_ph;                       /* W0_PH: phonon frequency */
  const double *bath_h;                /* replica only: Hbath(Nspin,Nspin,Norb,Norb,Nbath) = bath_from_sym(lambda) of every
                                        * replica, as ed_buildh_main assembles it (ED_HAMILTONIAN_SPARSE_HxV.f90:61-75) */
} edgpu_params;

typedef struct edgpu_ctx edgpu_ctx;    /* opaque: one per process, mirrors the module state */

/* ---- context ------------------------------------------------------------------------- */
/* Replaces the module-global setup consumed by build_Hv_sector (init_ed_structure,
 * ED_SETUP.f90:141-349).  device < 0: use the current CUDA device. */
int  edgpu_create(const edgpu_params *p, int device, edgpu_ctx **out);
int  edgpu_destroy(edgpu_ctx *c);
/* Update bath/interaction parameters between DMFT iterations (set_dmft_bath, ED_MAIN.f90:259-355). */
int  edgpu_set_params(edgpu_ctx *c, const edgpu_params *p);
const char *edgpu_last_error(void);
int  edgpu_set_option(edgpu_ctx *c, const char *key, int64_t value);
int  edgpu_device_count(int *n);

/* ---- communicator (replaces ed_set_MpiComm, ED_VARS_GLOBAL.f90:332-347) ---------------- */
/* One process per GPU.  Rank 0 calls edgpu_comm_unique_id, the host broadcasts the 128 bytes
 * (MPI_Bcast in the Fortran host, torch.distributed in the tests), every rank calls
 * edgpu_comm_init.  Uses NCCL over NVLink. */
int  edgpu_comm_unique_id(char id[128]);
int  edgpu_comm_init(edgpu_ctx *c, int rank, int nranks, const char id[128]);
int  edgpu_comm_finalize(edgpu_ctx *c);

/* ---- sector bookkeeping (ED_SETUP.f90:446-520) ------------------------------------------ */
int  edgpu_get_sector(const edgpu_ctx *c, int nup, int ndw, int *isector);     /* get_Sector */
int  edgpu_get_nup_ndw(const edgpu_ctx *c, int isector, int *nup, int *ndw);   /* get_Nup/get_Ndw */
/* Shard geometry, pure host arithmetic (ED_HAMILTONIAN.f90:96-110, ED_HAMILTONIAN_COMMON.f90:62-79):
 * q = n/P (+1 if r < mod(n,P)), off = first index owned by r. */
void edgpu_split(int64_t n, int nranks, int rank, int64_t *q, int64_t *off);
/* Send/receive offsets and counts (in doubles, arrays of nranks) of the grouped all-to-all that
 * replaces vector_transpose_MPI (ED_HAMILTONIAN_COMMON.f90:53-118); dir 0: V(DimUp,qdw) ->
 * Vt(DimDw,qup), dir 1: the way back.  Pure host arithmetic (block layouts: see comm.cu). */
void edgpu_transpose_plan(int64_t dimup, int64_t dimdw, int nranks, int rank, int dir,
                          int64_t *soff, int64_t *scnt, int64_t *roff, int64_t *rcnt);

/* ---- operator lifecycle ------------------------------------------------------------------ */
/* build_Hv_sector(isector), ED_HAMILTONIAN.f90:43-168: builds the basis maps on device
 * (build_sector, ED_SETUP.f90:745-780), the shard geometry, and -- when ed_sparse_h -- the
 * stored factors spH0d/spH0ups/spH0dws/spH0nd (ed_buildh_main,
 * ED_HAMILTONIAN_SPARSE_HxV.f90:25-200).  One live sector per context (Hstatus). */
int  edgpu_build_hv_sector(edgpu_ctx *c, int isector);
/* delete_Hv_sector, ED_HAMILTONIAN.f90:174-222 */
int  edgpu_delete_hv_sector(edgpu_ctx *c);
/* vecDim_Hv_sector, ED_HAMILTONIAN.f90:229-253 */
int  edgpu_vecdim_hv_sector(const edgpu_ctx *c, int isector, int64_t *vecdim);

/* ---- the operator ---------------------------------------------------------------------------- */
/* dd_sparse_HxV(Nloc,v,Hv) (ED_VARS_GLOBAL.f90:75-81) as held by spHtimesV_p: host in, host out.
 * Replaces spMatVec_main / spMatVec_MPI_main (ED_HAMILTONIAN_SPARSE_HxV.f90:391-485, 568-694)
 * and directMatVec_main / directMatVec_MPI_main (ED_HAMILTONIAN_DIRECT_HxV.f90:21-95, 180-284),
 * selected by ed_sparse_h like build_Hv_sector does (ED_HAMILTONIAN.f90:139-166).  With NPH > 0 the phonon
 * terms (:445-468) are included and nloc = DimUp*mpiQdw*DimPh; with ed_total_ud = 0 the operator is spMatVec_orbs
 * (:487-564).  On a sharded sector the call is collective (every rank applies the operator). */
int  edgpu_hxv(edgpu_ctx *c, int64_t nloc, const double *v, double *hv);
/* Fortran procedure-pointer compatible form: uses the context made current by the last
 * edgpu_build_hv_sector in this process; aborts (like the reference's stop) on error. */
void edgpu_sphtimesv(const int32_t *nloc, const double *v, double *hv);
/* Same operator on device pointers already resident in HBM (no copies). */
int  edgpu_hxv_device(edgpu_ctx *c, int64_t nloc, const double *d_v, double *d_hv);

/* ---- Lanczos (SciFortran call signatures used by ED_DIAG / ED_GF_NORMAL) ---------------------- */
/* sp_lanc_eigh([MpiComm,] MatVec, egs, vect, Nitermax, iverbose, threshold, ncheck)
 * (ED_DIAG.f90:174-186).  vect in: start vector (all zero => pseudo-random), out: eigenvector;
 * local shard of length nloc.  All Lanczos vectors stay on device.  Optional outputs
 * (may be NULL): nlanc, alanc[nitermax], blanc[nitermax] (blanc[0]=0, blanc[k]=b_k). */
int  edgpu_sp_lanc_eigh(edgpu_ctx *c, double *egs, double *vect, int64_t nloc, int nitermax,
                        int iverbose, double threshold, int ncheck,
                        int *nlanc, double *alanc, double *blanc);
/* sp_lanc_tridiag([MpiComm,] MatVec, vin, alanc, blanc) (ED_GF_NORMAL.f90:232-237): nlanc steps
 * from vin (destroyed in the reference; left untouched here). */
int  edgpu_sp_lanc_tridiag(edgpu_ctx *c, const double *vin, int64_t nloc, double *alanc,
                           double *blanc, int nlanc, double threshold);

/* ---- sector scan (ed_diag_d, ED_DIAG.f90:83-276) -------------------------------------------------------------------- */
/* Lowest eigenvalue of every listed sector: build_Hv_sector, Lanczos from the pseudo-random start vector with
 * Nitermax = min(dim, nitermax) (the reference diagonalises sectors below lanc_dim_threshold densely and the others
 * with sp_lanc_eigh), delete_Hv_sector.  twin != 0: a sector (Nup < Ndw) whose mirror (Ndw, Nup) is in the list is
 * not diagonalised again (ed_twin: same spectrum when both spins have the same parameters).  The eigenvector of the
 * lowest sector stays on the device as the state of the chains / observables (no host round trip); *best = its
 * position in the list.  e0[nsectors]; nlanc[nsectors] may be NULL. */
int  edgpu_diag_sectors(edgpu_ctx *c, int nsectors, const int *isector, int nitermax, double threshold,
                        int ncheck, int twin, double *e0, int *nlanc, int *best);

/* ---- ed_total_ud = F: one (Nup, Ndw) pair per orbital (Ns_Ud = Norb, Ns_Orb = 1 + Nbath) ------------------------
 * ed_buildh_orbs / spMatVec_orbs / directMatVec_orbs (ED_HAMILTONIAN_SPARSE_HxV.f90:206-370, 487-564) with
 * stored/Orbs/H_local.f90, H_up.f90, H_dw.f90; needs Jx = Jp = 0 (ED_SETUP.f90:69-71).  With params.ed_total_ud = 0
 * the sector number is get_Sector([Nups, Ndws], Ns_Orb) (ED_SETUP.f90:446-457); edgpu_build_hv_sector, edgpu_hxv*,
 * edgpu_sp_lanc_eigh, edgpu_sp_lanc_tridiag and edgpu_diag_sectors work on it unchanged (the vector runs over
 * [DimUps, DimDws], first index fastest, ED_SETUP.f90:520-545).  Single rank; chains and observables of such a state
 * return EDGPU_ERR_UNSUPPORTED. */
int  edgpu_get_sector_orbs(const edgpu_ctx *c, const int *nups, const int *ndws, int *isector);
int  edgpu_get_qn_orbs(const edgpu_ctx *c, int isector, int *nups, int *ndws);           /* get_Nup / get_Ndw, :477-500 */
/* introspection of the live sector: dims[2*Norb] (up words of orbital 1..Norb, then the dw words); factor f: map =
 * Hs(f)%map, CSR of spH0ups(f+1) resp. spH0dws(f+1-Norb) in insertion order (0-based; arrays may be NULL); the stored
 * diagonal spH0d (ed_sparse_h = 1). */
int  edgpu_get_orbs_dims(const edgpu_ctx *c, int64_t *dims, int64_t *dim);
int  edgpu_get_orbs_factor(const edgpu_ctx *c, int f, int32_t *map, int64_t *nnz, int64_t *rowptr, int64_t *cols, double *vals);
int  edgpu_get_orbs_diag(const edgpu_ctx *c, double *out);

/* ---- Green's function chains (lanc_build_gf_normal_main, ED_GF_NORMAL.f90:124-334) --------------- */
/* Keeps the ground state of sector (nup,ndw) on device for the chains; gs is the local shard.
 * Replaces es_return_cvector + the master-only c/cdg loops (:184-216, :259-290). */
int  edgpu_gf_set_state(edgpu_ctx *c, int isector, const double *gs, int64_t nloc, double e0);
/* The same hand-off without the host round trip: the eigenvector that the last edgpu_sp_lanc_eigh left on the
 * device (its sector must still be live) becomes the state of the chains; every rank keeps its own shard.
 * Replaces es_return_cvector's gather to the master rank (ED_EIGENSPACE.f90:502-572). */
int  edgpu_gf_set_state_from_eigh(edgpu_ctx *c);
/* Batched chains: for ch < nchains applies c^+ (addrem[ch]=+1) or c (addrem[ch]=-1) of orbital
 * iorb[ch] (1-based) and spin ispin[ch] (1=up,2=dw) to the stored state on device, normalises,
 * and runs nlanc_max Lanczos steps in the target sector (chains with the same target sector
 * share one build_Hv_sector).  On a sharded state a spin-up operator acts inside the local columns;
 * a spin-down operator is a signed permutation of i_dw columns and is applied through one grouped
 * column exchange between the ranks (no gather on a master rank).  Outputs per chain: norm2[ch], nlanc[ch] (=min(jdim,nlanc_max), 0 if
 * the target sector does not exist), alanc/blanc rows of length nlanc_max. */
int  edgpu_gf_chains(edgpu_ctx *c, int nchains, const int *iorb, const int *ispin,
                     const int *addrem, int nlanc_max, double threshold,
                     double *norm2, int *nlanc, double *alanc, double *blanc);
/* add_to_lanczos_gf_normal (ED_GF_NORMAL.f90:599-654), T=0: G(z_i) += norm2/zeta *
 * sum_j Z(1,j)^2 / (z_i - isign*(eps_j - ei)); z and g are interleaved (re,im) pairs.  Host-side
 * O(nlanc*L) work on the chain output; no vector is touched. */
int  edgpu_add_to_lanczos_gf(double norm2, double zeta, double ei, const double *alanc,
                             const double *blanc, int nlanc, int isign,
                             const double *z, int nz, double *g);
/* Susceptibility chains: lanc_ed_build_spinChi_main / _tot_main / _mix_main (ED_GF_CHISPIN.f90:114-415) for kind = 0
 * and lanc_ed_build_densChi_* (ED_GF_CHIDENS.f90:111-426) for kind = 1, ed_total_ud = T.  Start vector O|state> with
 * the diagonal operator O = Sz resp. n of orbital iorb (iorb == jorb), of all impurity orbitals (iorb == 0) or the mixed
 * combination O_iorb + O_jorb, applied on the device to the state of edgpu_gf_set_state* in its OWN sector (each rank
 * scales its shard; the reference builds it on the master and scatters), then sp_lanc_tridiag; the channels run as a
 * batch.  Outputs as edgpu_gf_chains. */
int  edgpu_chi_chains(edgpu_ctx *c, int kind, int nchains, const int *iorb, const int *jorb, int nlanc_max,
                      double threshold, double *norm2, int *nlanc, double *alanc, double *blanc);
/* add_to_lanczos_spinChi (ED_GF_CHISPIN.f90:434-488) == add_to_lanczos_densChi (ED_GF_CHIDENS.f90:436-489), T = 0:
 * accumulates chi_iv[0..lmats] on the bosonic frequencies vm[0..lmats], chi_tau[0..ltau] on tau[0..ltau] and
 * chi_w[lreal] (interleaved complex) on vr[lreal] + i eps; any output may be NULL. */
int  edgpu_add_to_lanczos_chi(double norm2, double zeta, double ei, double beta, const double *alanc, const double *blanc,
                              int nlanc, const double *vm, int lmats, double *chi_iv, const double *tau, int ltau,
                              double *chi_tau, const double *vr, int lreal, double eps, double *chi_w);

/* ---- observables (lanc_observables, ED_OBSERVABLES.f90:95-363; lanc_local_energy, :372-600) ------------------- */
/* Local observables and energies of the state kept on the device for the chains (edgpu_gf_set_state /
 * edgpu_gf_set_state_from_eigh), T = 0, one state, zeta = zeta_function (number of degenerate ground states).
 * The state is not gathered: every rank reduces its own shard and the 4^Norb weights are all-reduced, so the
 * call is collective and every rank gets the same numbers (the reference broadcasts them from the master).
 * Arrays are indexed like the reference's with leading dimension EDGPU_MAX_ORB: sz2[iorb + 5*jorb] = sz2(iorb,jorb),
 * dm[ispin][iorb + 5*jorb] = imp_density_matrix(ispin,ispin,iorb,jorb) for ispin <= Nspin, prob[i] = Prob(i+1);
 * epot includes ehartree (ED_OBSERVABLES.f90:587). */
typedef struct edgpu_observables {
  double dens[EDGPU_MAX_ORB], dens_up[EDGPU_MAX_ORB], dens_dw[EDGPU_MAX_ORB], docc[EDGPU_MAX_ORB], magz[EDGPU_MAX_ORB];
  double sz2[EDGPU_MAX_ORB * EDGPU_MAX_ORB], n2[EDGPU_MAX_ORB * EDGPU_MAX_ORB];
  double s2tot;
  double prob[243];
  double dm[2][EDGPU_MAX_ORB * EDGPU_MAX_ORB];
  double eknot, epot, ehartree, dust, dund, dse, dph;
} edgpu_observables;
int  edgpu_observables_normal(edgpu_ctx *c, double zeta, edgpu_observables *out);

/* ---- introspection for bit-exact checks (Appendix C of SURVEY.md) ---------------------------------- */
int  edgpu_get_dims(const edgpu_ctx *c, int64_t *dimup, int64_t *dimdw, int64_t *qdw,
                    int64_t *ishift, int64_t *nloc);
/* which: 0 = Hs(1)%map (up), 1 = Hs(2)%map (dw); out has DimUp / DimDw int32 entries */
int  edgpu_get_sector_map(const edgpu_ctx *c, int which, int32_t *out);
/* which: 0 = spH0ups(1), 1 = spH0dws(1), 2 = spH0nd.  Call with rowptr=NULL to query sizes.
 * CSR in the reference's insertion order, 0-based, int64. */
int  edgpu_get_csr(const edgpu_ctx *c, int which, int64_t *nrow, int64_t *nnz,
                   int64_t *rowptr, int64_t *cols, double *vals);
/* spH0d, one value per local ELECTRON row (nloc = DimUp*mpiQdw; the phonon slabs share it): stored mode, or the
 * recomputed diagonal (direct mode) */
int  edgpu_get_diag(const edgpu_ctx *c, double *out, int64_t nloc);

/* ---- device-resident vector helpers (used by bench.py and by hosts that keep vectors in HBM) ------ */
int  edgpu_dev_alloc(edgpu_ctx *c, int64_t nbytes, void **dptr);
int  edgpu_dev_free(edgpu_ctx *c, void *dptr);
int  edgpu_dev_upload(edgpu_ctx *c, void *dptr, const void *host, int64_t nbytes);
int  edgpu_dev_download(edgpu_ctx *c, void *host, const void *dptr, int64_t nbytes);
int  edgpu_dev_fill_bench_vector(edgpu_ctx *c, double *d_v, int64_t nloc, int64_t global_offset);
int  edgpu_sync(edgpu_ctx *c);
/* <a, b> of two device-resident sector vectors (local shards of length nloc; the partial sums are all-reduced over
 * the ranks, so the call is collective): dot_product + MPI_Allreduce of the reference's MPI Lanczos. */
int  edgpu_dev_dot(edgpu_ctx *c, int64_t nloc, const double *d_a, const double *d_b, double *out);
/* Times `reps` back-to-back device-resident H*v with CUDA events on the engine's stream;
 * ms_total = elapsed over all reps.  ms_kernel[k] (k < 8, may be NULL) = per-kernel-class totals. */
int  edgpu_time_hxv_device(edgpu_ctx *c, int64_t nloc, const double *d_v, double *d_hv,
                           int reps, double *ms_total);
/* Per-kernel split of the same loop (any rank count): ms_pass[i] = total over reps of pass i (i < *npasses <= 6),
 * names = 6 x 32 bytes of NUL-terminated kernel names.  Used for bench.py's roofline of the dominant kernel. */
int  edgpu_time_hxv_passes(edgpu_ctx *c, int64_t nloc, const double *d_v, double *d_hv, int reps,
                           int *npasses, double *ms_pass, char *names);
/* One device-resident Lanczos iteration loop (reps steps, no host sync inside), for iter/s. */
int  edgpu_time_lanczos_device(edgpu_ctx *c, int64_t nloc, double *d_v0, int reps, double *ms_total);
/* Halo of the sharded fast path of the live sector (ED_HAMILTONIAN_COMMON.f90:53-118 replaced): bytes this rank
 * stores into its peers' halo buffers per H*v, bytes it receives, column windows of the pipeline.  All zero when the
 * sector does not use the halo path (one rank, multi-orbital model, no_peer). */
int  edgpu_halo_info(edgpu_ctx *c, int64_t *bytes_out, int64_t *bytes_in, int *windows);
/* Kernel launches issued by this context since creation (bench.py's gpu_launches). */
int  edgpu_launch_count(const edgpu_ctx *c, int64_t *n);

#ifdef __cplusplus
}
#endif
#endif
