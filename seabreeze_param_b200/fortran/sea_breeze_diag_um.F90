!------------------------------------------------------------------------------
! sea_breeze_diag_mod for the UM hook -- ISO_C_BINDING shim over include/sbd.h.
!
! Drop-in for UM/vn10.7/sea_breeze_diag.F90 (module sea_breeze_diag_mod): the same
! four routines with the UM's own dummy lists, array bounds taken from
! atm_fields_bounds_mod exactly as the reference declares them:
!   seabreeze_diag(timestep, timestep_number, p, u, v, theta, z, sigma, mask,       U:55-56
!                  windspeed, winddir, thc, sb_con, error)
!       p, u, v: pdims;  theta, z, sigma: tdims_s;  mask: tdims_l;                  U:78-117
!       windspeed, winddir, thc: tdims;  sb_con: pdims;  error out (U:102, 198-201);
!       theta is overwritten by t0 = theta - gmma z sigmoid(sigma)                  U:210-211
!   get_edges(mask, icefrac, landfrac)            mask: tdims_l, icefrac(:,:,1)      U:328-341
!   get_dist(landfrac, coast)                     tdims_l; true_latitude/longitude   U:448-476
!   sigmoid(ary, sm)                              tdims_s                            U:603
! and the UM's 8-argument swap_bounds (module halo_exchange_sbd below, U:408-410) for a
! single domain.  All arithmetic is in csrc/sbd_cuda.cu (sbd_um_* entry points).
!
! REAL kind: the UM is normally built with promoted reals (real = real(kind=jprb), 64-bit).
! Compile this file with -DSBD_R8 in that case: the 64-bit entry points are used.
! Without it default REAL must be 32-bit.
!
! NOTE: no Fortran compiler (and no UM) exists in the build image of this repository: this
! file is compiled by the integrator.  The same entry points are exercised through
! seabreeze_param_b200/um.py by tests/test_gpu_um.py.
!------------------------------------------------------------------------------
#if defined(SBD_R8)
#define SBD_REAL c_double
#define SBD_UM_DIAG sbd_um_seabreeze_diag_f64
#define SBD_UM_EDGES sbd_um_get_edges_f64
#define SBD_UM_DIST sbd_um_get_dist_f64
#define SBD_UM_SWAP sbd_um_swap_bounds_f64
#define SBD_SIGMOID sbd_sigmoid_f64
#else
#define SBD_REAL c_float
#define SBD_UM_DIAG sbd_um_seabreeze_diag_f32
#define SBD_UM_EDGES sbd_um_get_edges_f32
#define SBD_UM_DIST sbd_um_get_dist_f32
#define SBD_UM_SWAP sbd_um_swap_bounds_f32
#define SBD_SIGMOID sbd_sigmoid_f32
#endif

module sbd_um_c_api
  use, intrinsic :: iso_c_binding, only : c_int, c_float, c_double, c_ptr
  implicit none
  interface
    function SBD_UM_DIAG(timestep, timestep_number, p, u, v, theta, z, sigma, mask, windspeed, winddir, &
                         thc, sb_con, row_length, rows, levels, halo_s, halo_l, theta_to_t0, prm, flags) &
        bind(C) result(rc)
      import :: c_int, SBD_REAL, c_ptr
      real(SBD_REAL), value :: timestep
      integer(c_int), value :: timestep_number
      real(SBD_REAL), intent(in) :: p(*), u(*), v(*), z(*), sigma(*), mask(*)
      real(SBD_REAL), intent(inout) :: theta(*), windspeed(*), winddir(*), thc(*), sb_con(*)
      integer(c_int), value :: row_length, rows, levels, halo_s, halo_l, theta_to_t0
      type(c_ptr), value :: prm, flags
      integer(c_int) :: rc
    end function
    function SBD_UM_EDGES(mask, icefrac, landfrac, row_length, rows, halo_i, halo_j) bind(C) result(rc)
      import :: c_int, SBD_REAL
      real(SBD_REAL), intent(inout) :: mask(*)
      real(SBD_REAL), intent(in) :: icefrac(*), landfrac(*)
      integer(c_int), value :: row_length, rows, halo_i, halo_j
      integer(c_int) :: rc
    end function
    function SBD_UM_DIST(landfrac, coast, true_latitude, true_longitude, row_length, rows, halo_i, halo_j, &
                         maxdist) bind(C) result(rc)
      import :: c_int, SBD_REAL
      real(SBD_REAL), intent(in) :: landfrac(*), true_latitude(*), true_longitude(*)
      real(SBD_REAL), intent(inout) :: coast(*)
      integer(c_int), value :: row_length, rows, halo_i, halo_j
      real(SBD_REAL), value :: maxdist
      integer(c_int) :: rc
    end function
    function SBD_UM_SWAP(field, row_len, rows, levels, halo_i, halo_j) bind(C) result(rc)
      import :: c_int, SBD_REAL
      real(SBD_REAL), intent(inout) :: field(*)
      integer(c_int), value :: row_len, rows, levels, halo_i, halo_j
      integer(c_int) :: rc
    end function
    function SBD_SIGMOID(ary, sm, nx, ny, stats) bind(C) result(rc)
      import :: c_int, SBD_REAL, c_ptr
      real(SBD_REAL), intent(in) :: ary(*)
      real(SBD_REAL), intent(out) :: sm(*)
      integer(c_int), value :: nx, ny
#if !defined(SBD_R8)
      type(c_ptr), value :: stats
#endif
      integer(c_int) :: rc
    end function
  end interface
end module sbd_um_c_api

! The UM's halo swap for one domain (U:408-410 call it through `use halo_exchange, only: swap_bounds`; give this
! module that name, or keep the UM's own MPI version when the run is decomposed over PEs).
module halo_exchange_sbd
  use, intrinsic :: iso_c_binding, only : c_int
  use sbd_um_c_api
  implicit none
contains
  subroutine swap_bounds(field, row_len, rows, levels, halo_i, halo_j, fld_type, l_vector)
    integer, intent(in) :: row_len, rows, levels, halo_i, halo_j, fld_type
    logical, intent(in) :: l_vector
    real, intent(inout) :: field(1 - halo_i:row_len + halo_i, 1 - halo_j:rows + halo_j, levels)
    integer(c_int) :: rc
    rc = SBD_UM_SWAP(field, row_len, rows, levels, halo_i, halo_j)
    if (rc /= 0) error stop 'halo_exchange_sbd: swap_bounds failed'
  end subroutine swap_bounds
end module halo_exchange_sbd

module sea_breeze_diag_mod
  use, intrinsic :: iso_c_binding, only : c_int, c_null_ptr
  use sbd_um_c_api
  implicit none
  private
  public :: seabreeze_diag, get_edges, get_dist, sigmoid
  character(len=*), parameter, private :: ModuleName = 'sea_breeze_diag_mod'
  real, parameter :: maxdist = 180.      ! U:168
contains

  subroutine seabreeze_diag(timestep, timestep_number, &
      p, u, v, theta, z, sigma, mask, windspeed, winddir, thc, sb_con, error)
    use atm_fields_bounds_mod, only : pdims, tdims, tdims_s, tdims_l
    integer, intent(in) :: timestep_number
    real, intent(in) :: timestep
    real, intent(in) :: p(pdims%i_start:pdims%i_end, pdims%j_start:pdims%j_end, pdims%k_start:pdims%k_end)
    real, intent(in) :: u(pdims%i_start:pdims%i_end, pdims%j_start:pdims%j_end, pdims%k_start:pdims%k_end)
    real, intent(in) :: v(pdims%i_start:pdims%i_end, pdims%j_start:pdims%j_end, pdims%k_start:pdims%k_end)
    real, intent(in) :: z(tdims_s%i_start:tdims_s%i_end, tdims_s%j_start:tdims_s%j_end)
    real, intent(in) :: sigma(tdims_s%i_start:tdims_s%i_end, tdims_s%j_start:tdims_s%j_end)
    real, intent(in) :: mask(tdims_l%i_start:tdims_l%i_end, tdims_l%j_start:tdims_l%j_end)
    integer, intent(out) :: error
    real, intent(inout) :: sb_con(pdims%i_start:pdims%i_end, pdims%j_start:pdims%j_end)
    real, intent(inout) :: windspeed(tdims%i_start:tdims%i_end, tdims%j_start:tdims%j_end)
    real, intent(inout) :: winddir(tdims%i_start:tdims%i_end, tdims%j_start:tdims%j_end)
    real, intent(inout) :: thc(tdims%i_start:tdims%i_end, tdims%j_start:tdims%j_end)
    real, intent(inout) :: theta(tdims_s%i_start:tdims_s%i_end, tdims_s%j_start:tdims_s%j_end)
    integer(c_int) :: rc, row_length, rows, levels
    row_length = pdims%i_end - pdims%i_start + 1
    rows = pdims%j_end - pdims%j_start + 1
    levels = pdims%k_end - pdims%k_start + 1
    ! the two halo kinds must be square for the window search (the UM's tdims_s has halo_i == halo_j == 1 by
    ! default; runs that use this diagnostic extend it): the narrower direction bounds the search
    rc = SBD_UM_DIAG(timestep, timestep_number, p, u, v, theta, z, sigma, mask, windspeed, winddir, thc, sb_con, &
                     row_length, rows, levels, min(tdims_s%halo_i, tdims_s%halo_j), min(tdims_l%halo_i, tdims_l%halo_j), &
                     1_c_int, c_null_ptr, c_null_ptr)
    error = rc                       ! 0 ok, 1 = bad grid/subset definition (U:198-201), > 1 = device-side failure
  end subroutine seabreeze_diag

  subroutine get_edges(mask, icefrac, landfrac)
    use atm_fields_bounds_mod, only : tdims, tdims_l
    real, intent(out) :: mask(tdims_l%i_start:tdims_l%i_end, tdims_l%j_start:tdims_l%j_end)
    real, intent(in) :: landfrac(tdims%i_start:tdims%i_end, tdims%j_start:tdims%j_end)
    real, intent(in) :: icefrac(tdims%i_start:tdims%i_end, tdims%j_start:tdims%j_end, 1)
    integer(c_int) :: rc
    rc = SBD_UM_EDGES(mask, icefrac, landfrac, tdims%i_end - tdims%i_start + 1, tdims%j_end - tdims%j_start + 1, &
                      tdims_l%halo_i, tdims_l%halo_j)
    if (rc /= 0) error stop 'sea_breeze_diag_mod: get_edges failed'
  end subroutine get_edges

  subroutine get_dist(landfrac, coast)
    use trignometric_mod, only : true_latitude, true_longitude
    use atm_fields_bounds_mod, only : tdims, tdims_l
    real, intent(in) :: landfrac(tdims_l%i_start:tdims_l%i_end, tdims_l%j_start:tdims_l%j_end)
    real, intent(inout) :: coast(tdims_l%i_start:tdims_l%i_end, tdims_l%j_start:tdims_l%j_end)
    integer(c_int) :: rc
    rc = SBD_UM_DIST(landfrac, coast, true_latitude, true_longitude, tdims%i_end - tdims%i_start + 1, &
                     tdims%j_end - tdims%j_start + 1, tdims_l%halo_i, tdims_l%halo_j, maxdist)
    if (rc /= 0) error stop 'sea_breeze_diag_mod: get_dist failed'
  end subroutine get_dist

  subroutine sigmoid(ary, sm)
    use atm_fields_bounds_mod, only : tdims_s
    real, intent(in) :: ary(tdims_s%i_start:tdims_s%i_end, tdims_s%j_start:tdims_s%j_end)
    real, intent(out) :: sm(tdims_s%i_start:tdims_s%i_end, tdims_s%j_start:tdims_s%j_end)
    integer(c_int) :: rc
#if defined(SBD_R8)
    rc = SBD_SIGMOID(ary, sm, tdims_s%i_end - tdims_s%i_start + 1, tdims_s%j_end - tdims_s%j_start + 1)
#else
    rc = SBD_SIGMOID(ary, sm, tdims_s%i_end - tdims_s%i_start + 1, tdims_s%j_end - tdims_s%j_start + 1, c_null_ptr)
#endif
    if (rc /= 0) error stop 'sea_breeze_diag_mod: sigmoid failed'
  end subroutine sigmoid

end module sea_breeze_diag_mod
