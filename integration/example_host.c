/* example_host.c -- the C-ABI of include/edgpu.h driven from plain C (no Python, no torch): the calls the Fortran shim
 * of ED_GPU_BINDINGS.f90 makes, in the order ed_diag_d / build_gf_normal make them for one sector.
 *   gcc -std=c99 -I include integration/example_host.c -L dmft-lanc-ed_b200 -ledgpu -Wl,-rpath,$PWD/dmft-lanc-ed_b200 -lm -o example_host
 * Needs a B200 to run (the engine has no CPU fallback); tests/test_abi.py compiles and links it on the CPU box. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "edgpu.h"

#define CHECK(call)                                                                  \
  do {                                                                               \
    int rc_ = (call);                                                                \
    if (rc_) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, edgpu_last_error()); return 1; } \
  } while (0)

int main(void) {
  /* single-band model, Nbath = 7 (config C1 of BASELINE.json): e_k = -2 .. 2, V = 1/sqrt(Nbath), U = 2, HFMODE */
  enum { NBATH = 7 };
  double e[NBATH], v[NBATH];
  for (int k = 0; k < NBATH; k++) { e[k] = -2.0 + 4.0 * k / (NBATH - 1); v[k] = 1.0 / sqrt((double)NBATH); }
  edgpu_params p = {0};
  p.norb = 1; p.nbath = NBATH; p.nspin = 1; p.hfmode = 1; p.ed_sparse_h = 0; p.nph = 0; p.ed_total_ud = 1; p.bath_type = 0;
  p.uloc[0] = 2.0;
  p.bath_e = e; p.bath_v = v;
  edgpu_ctx *ctx = NULL;
  CHECK(edgpu_create(&p, -1, &ctx));
  int isector = 0;
  CHECK(edgpu_get_sector(ctx, 4, 4, &isector));
  CHECK(edgpu_build_hv_sector(ctx, isector));                      /* build_Hv_sector(isector) */
  int64_t nloc = 0;
  CHECK(edgpu_vecdim_hv_sector(ctx, isector, &nloc));              /* vecDim_Hv_sector */
  double *vec = (double *)calloc((size_t)nloc, sizeof(double)), *hv = (double *)calloc((size_t)nloc, sizeof(double));
  for (int64_t i = 0; i < nloc; i++) vec[i] = 1.0 / sqrt((double)nloc);
  int32_t n32 = (int32_t)nloc;
  edgpu_sphtimesv(&n32, vec, hv);                                  /* spHtimesV_p(Nloc, v, Hv) */
  double egs = 0.0;
  int nlanc = 0;
  CHECK(edgpu_sp_lanc_eigh(ctx, &egs, vec, nloc, 512, 0, 1e-18, 10, &nlanc, NULL, NULL));   /* sp_lanc_eigh */
  printf("E0 = %.12f after %d Lanczos steps (expected -9.361735245469)\n", egs, nlanc);
  CHECK(edgpu_gf_set_state_from_eigh(ctx));                        /* the eigenvector stays on the device */
  CHECK(edgpu_delete_hv_sector(ctx));                              /* delete_Hv_sector() */
  int iorb[2] = {1, 1}, ispin[2] = {1, 1}, addrem[2] = {1, -1}, nl[2];
  double norm2[2], *a = (double *)calloc(400, sizeof(double)), *b = (double *)calloc(400, sizeof(double));
  CHECK(edgpu_gf_chains(ctx, 2, iorb, ispin, addrem, 200, 1e-12, norm2, nl, a, b));          /* lanc_build_gf_normal_main */
  printf("norm2(add) + norm2(remove) = %.12f (sum rule: 1)\n", norm2[0] + norm2[1]);
  edgpu_observables obs;
  CHECK(edgpu_observables_normal(ctx, 1.0, &obs));                 /* lanc_observables */
  printf("<n> = %.10f, <n_up n_dw> = %.10f\n", obs.dens[0], obs.docc[0]);
  free(vec); free(hv); free(a); free(b);
  CHECK(edgpu_destroy(ctx));
  return 0;
}
