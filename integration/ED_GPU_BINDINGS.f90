! ED_GPU_BINDINGS.f90 -- ISO_C_BINDING interface of the B200 engine (include/edgpu.h) for dmft-lanc-ed.
! The same text as INTEGRATION.md section 2, as a file a maintainer can add to the reference's source list
! (CMakeLists.txt: next to ED_VARS_GLOBAL.f90).  Not compiled in this repository: the image has no Fortran compiler.
module ED_GPU_BINDINGS
  use, intrinsic :: iso_c_binding
  implicit none
  type, bind(C) :: edgpu_params            ! include/edgpu.h: struct edgpu_params
     integer(c_int32_t) :: norb, nbath, nspin, hfmode, ed_sparse_h, nph, ed_total_ud, bath_type   ! bath_type: 0 normal, 1 hybrid, 2 replica
     real(c_double)     :: uloc(5), ust, jh, jx, jp, xmu
     type(c_ptr)        :: imphloc, bath_e, bath_v   ! impHloc(Nspin,Nspin,Norb,Norb), dmft_bath%e/v(Nspin,Norb,Nbath)
     real(c_double)     :: g_ph(5), w0_ph            ! G_PH, W0_PH (used when nph > 0)
     type(c_ptr)        :: bath_h                    ! replica: Hbath_tmp(Nspin,Nspin,Norb,Norb,Nbath) of ed_buildh_main, else c_null_ptr
  end type
  type, bind(C) :: edgpu_observables       ! include/edgpu.h: struct edgpu_observables (leading dimension 5 = EDGPU_MAX_ORB)
     real(c_double) :: dens(5), dens_up(5), dens_dw(5), docc(5), magz(5), sz2(5,5), n2(5,5), s2tot, prob(243)
     real(c_double) :: dm(5,5,2)            ! dm(iorb,jorb,ispin) = imp_density_matrix(ispin,ispin,iorb,jorb)
     real(c_double) :: eknot, epot, ehartree, dust, dund, dse, dph
  end type
  type(c_ptr), save :: ed_gpu_ctx = c_null_ptr
  interface
     integer(c_int) function edgpu_create(p, device, ctx) bind(C, name="edgpu_create")
       import; type(edgpu_params), intent(in) :: p; integer(c_int), value :: device; type(c_ptr) :: ctx
     end function
     integer(c_int) function edgpu_set_params(ctx, p) bind(C, name="edgpu_set_params")
       import; type(c_ptr), value :: ctx; type(edgpu_params), intent(in) :: p
     end function
     integer(c_int) function edgpu_comm_unique_id(id) bind(C, name="edgpu_comm_unique_id")
       import; character(kind=c_char) :: id(128)
     end function
     integer(c_int) function edgpu_comm_init(ctx, rank, nranks, id) bind(C, name="edgpu_comm_init")
       import; type(c_ptr), value :: ctx; integer(c_int), value :: rank, nranks; character(kind=c_char) :: id(128)
     end function
     integer(c_int) function edgpu_build_hv_sector(ctx, isector) bind(C, name="edgpu_build_hv_sector")
       import; type(c_ptr), value :: ctx; integer(c_int), value :: isector
     end function
     integer(c_int) function edgpu_delete_hv_sector(ctx) bind(C, name="edgpu_delete_hv_sector")
       import; type(c_ptr), value :: ctx
     end function
     subroutine edgpu_sphtimesv(nloc, v, hv) bind(C, name="edgpu_sphtimesv")   ! dd_sparse_HxV, ED_VARS_GLOBAL.f90:75-81
       import; integer(c_int32_t) :: nloc; real(c_double) :: v(*), hv(*)
     end subroutine
     integer(c_int) function edgpu_sp_lanc_eigh(ctx, egs, vect, nloc, nitermax, iverbose, threshold, ncheck, &
                                                nlanc, alanc, blanc) bind(C, name="edgpu_sp_lanc_eigh")
       import; type(c_ptr), value :: ctx; real(c_double) :: egs, vect(*); integer(c_int64_t), value :: nloc
       integer(c_int), value :: nitermax, iverbose, ncheck; real(c_double), value :: threshold
       type(c_ptr), value :: nlanc, alanc, blanc
     end function
     integer(c_int) function edgpu_sp_lanc_tridiag(ctx, vin, nloc, alanc, blanc, nlanc, threshold) &
                                                   bind(C, name="edgpu_sp_lanc_tridiag")
       import; type(c_ptr), value :: ctx; real(c_double) :: vin(*), alanc(*), blanc(*)
       integer(c_int64_t), value :: nloc; integer(c_int), value :: nlanc; real(c_double), value :: threshold
     end function
     integer(c_int) function edgpu_gf_set_state(ctx, isector, gs, nloc, e0) bind(C, name="edgpu_gf_set_state")
       import; type(c_ptr), value :: ctx; integer(c_int), value :: isector; real(c_double) :: gs(*)
       integer(c_int64_t), value :: nloc; real(c_double), value :: e0
     end function
     integer(c_int) function edgpu_gf_chains(ctx, nchains, iorb, ispin, addrem, nlanc_max, threshold, norm2, nlanc, &
                                             alanc, blanc) bind(C, name="edgpu_gf_chains")
       import; type(c_ptr), value :: ctx; integer(c_int), value :: nchains, nlanc_max; real(c_double), value :: threshold
       integer(c_int) :: iorb(*), ispin(*), addrem(*), nlanc(*); real(c_double) :: norm2(*), alanc(*), blanc(*)
     end function
     integer(c_int) function edgpu_gf_set_state_from_eigh(ctx) bind(C, name="edgpu_gf_set_state_from_eigh")
       import; type(c_ptr), value :: ctx
     end function
     integer(c_int) function edgpu_chi_chains(ctx, kind, nchains, iorb, jorb, nlanc_max, threshold, norm2, nlanc, &
                                              alanc, blanc) bind(C, name="edgpu_chi_chains")   ! ED_GF_CHISPIN / ED_GF_CHIDENS
       import; type(c_ptr), value :: ctx; integer(c_int), value :: kind, nchains, nlanc_max; real(c_double), value :: threshold
       integer(c_int) :: iorb(*), jorb(*), nlanc(*); real(c_double) :: norm2(*), alanc(*), blanc(*)
     end function
     integer(c_int) function edgpu_diag_sectors(ctx, nsectors, isector, nitermax, threshold, ncheck, twin, e0, nlanc, best) &
                                                bind(C, name="edgpu_diag_sectors")               ! ed_diag_d, ED_DIAG.f90:83-276
       import; type(c_ptr), value :: ctx; integer(c_int), value :: nsectors, nitermax, ncheck, twin
       real(c_double), value :: threshold; integer(c_int) :: isector(*), nlanc(*), best; real(c_double) :: e0(*)
     end function
     integer(c_int) function edgpu_observables_normal(ctx, zeta, obs) bind(C, name="edgpu_observables_normal")
       import; type(c_ptr), value :: ctx; real(c_double), value :: zeta; type(edgpu_observables) :: obs
     end function
     function edgpu_last_error() bind(C, name="edgpu_last_error") result(p)
       import; type(c_ptr) :: p
     end function
  end interface
contains
  subroutine edgpu_check(rc)
    integer(c_int), intent(in) :: rc
    if (rc /= 0) stop "edgpu ERROR (see edgpu_last_error)"
  end subroutine
  ! the procedure the pointer is set to: abstract interface dd_sparse_HxV is not bind(C), hence this shim
  subroutine gpuMatVec(Nloc, v, Hv)
    integer :: Nloc
    real(8), dimension(Nloc) :: v, Hv
    call edgpu_sphtimesv(int(Nloc, c_int32_t), v, Hv)
  end subroutine
end module
