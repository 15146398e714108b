!------------------------------------------------------------------------------
! sea_breeze_diag_mod -- ISO_C_BINDING shim over the CUDA C-ABI (include/sbd.h).
!
! Same module name and the same four public routines as the reference's
! generic/sea_breeze_diag.f90 (module sea_breeze_diag_mod, G:19-25), so a host
! model that says
!     use sea_breeze_diag_mod, only : get_edges, get_dist, seabreeze_diag
! (generic/dummy_model.f90:28,41) links against this file + libsbd_b200.so
! instead of the Fortran kernels.  Every routine only takes sizes, infers the
! halo width and forwards the arrays: all arithmetic is in csrc/sbd_cuda.cu.
!
! Argument order of seabreeze_diag: the dummy-argument NAMES are those of
! G:55-56; the ORDER is the one its callers use (dummy_model.f90:52-54 and the
! UM hook UM/vn10.7/sea_breeze_diag.F90:55-56): ..., theta, z, sigma, mask, ...
! Keyword calls are order-agnostic either way (SURVEY.md A.6).
!
! Default REAL must be 32-bit (c_float).  A UM build with promoted reals (-r8)
! uses sea_breeze_diag_mod_r8.f90 (the _f64 twins of the C-ABI) instead of this file.
!
! NOTE: no Fortran compiler exists in the build image of this repository, so
! this file is compiled by the integrator, not by the test-suite; the same
! call sequence is exercised through the C-ABI by csrc/dummy_model_harness.cpp.
!------------------------------------------------------------------------------
module sea_breeze_diag_mod
  use, intrinsic :: iso_c_binding, only : c_int, c_float, c_ptr, c_null_ptr
  implicit none
  private
  public :: seabreeze_diag, get_edges, get_dist, sigmoid

  interface
    function sbd_get_edges_f32(mask, icefrac, landfrac, nx, ny, h, swap, m1) &
        bind(C, name="sbd_get_edges_f32") result(rc)
      import :: c_int, c_float, c_ptr
      real(c_float), intent(inout) :: mask(*)
      real(c_float), intent(in) :: icefrac(*), landfrac(*)
      integer(c_int), value :: nx, ny, h, swap
      type(c_ptr), value :: m1
      integer(c_int) :: rc
    end function sbd_get_edges_f32

    function sbd_get_dist_f32(coast, landfrac, lon, lat, maxdist, cdist, nx, ny, h) &
        bind(C, name="sbd_get_dist_f32") result(rc)
      import :: c_int, c_float
      real(c_float), intent(in) :: coast(*), landfrac(*), lon(*), lat(*)
      real(c_float), value :: maxdist
      real(c_float), intent(out) :: cdist(*)
      integer(c_int), value :: nx, ny, h
      integer(c_int) :: rc
    end function sbd_get_dist_f32

    function sbd_sigmoid_f32(ary, sm, nx, ny, stats) bind(C, name="sbd_sigmoid_f32") result(rc)
      import :: c_int, c_float, c_ptr
      real(c_float), intent(in) :: ary(*)
      real(c_float), intent(out) :: sm(*)
      integer(c_int), value :: nx, ny
      type(c_ptr), value :: stats
      integer(c_int) :: rc
    end function sbd_sigmoid_f32

    function sbd_seabreeze_diag_f32(timestep, timestep_number, p, u, v, theta, mask, z, sigma, &
        windspeed, winddir, thc, sb_con, nx, ny, nz, h, prm, flags) &
        bind(C, name="sbd_seabreeze_diag_f32") result(rc)
      import :: c_int, c_float, c_ptr
      real(c_float), value :: timestep
      integer(c_int), value :: timestep_number
      real(c_float), intent(in) :: p(*), u(*), v(*), theta(*), mask(*), z(*), sigma(*)
      real(c_float), intent(inout) :: windspeed(*), winddir(*), thc(*), sb_con(*)
      integer(c_int), value :: nx, ny, nz, h
      type(c_ptr), value :: prm, flags
      integer(c_int) :: rc
    end function sbd_seabreeze_diag_f32
  end interface

contains

  subroutine check(rc, what)
    integer(c_int), intent(in) :: rc
    character(len=*), intent(in) :: what
    if (rc /= 0) then
      write (*, '(a,a,a,i0)') 'sea_breeze_diag_mod: ', what, ' failed, status ', rc
      error stop 1
    end if
  end subroutine check

  ! generic/sea_breeze_diag.f90:55-56 (order of the callers, see header)
  subroutine seabreeze_diag(timestep, timestep_number, p, u, v, theta, z, sigma, mask, &
                            windspeed, winddir, thc, sb_con)
    real, intent(in) :: timestep
    integer, intent(in) :: timestep_number
    real, intent(in), contiguous :: p(:, :, :), u(:, :, :), v(:, :, :)
    real, intent(in), contiguous :: theta(:, :), z(:, :), sigma(:, :), mask(:, :)
    real, intent(inout), contiguous :: windspeed(:, :), winddir(:, :), thc(:, :), sb_con(:, :)
    integer(c_int) :: nx, ny, nz, h
    nx = size(p, 1); ny = size(p, 2); nz = size(p, 3)
    h = (size(mask, 1) - nx) / 2          ! halo'd neighbour fields (UM tdims_s / tdims_l)
    call check(sbd_seabreeze_diag_f32(timestep, timestep_number, p, u, v, theta, mask, z, sigma, &
               windspeed, winddir, thc, sb_con, nx, ny, nz, h, c_null_ptr, c_null_ptr), 'seabreeze_diag')
  end subroutine seabreeze_diag

  ! generic/sea_breeze_diag.f90:273
  subroutine get_edges(mask, icefrac, landfrac, halo_size)
    real, intent(inout), contiguous :: mask(:, :)
    real, intent(in), contiguous :: icefrac(:, :), landfrac(:, :)
    integer, intent(in) :: halo_size
    call check(sbd_get_edges_f32(mask, icefrac, landfrac, size(landfrac, 1), size(landfrac, 2), &
               halo_size, 1_c_int, c_null_ptr), 'get_edges')
  end subroutine get_edges

  ! generic/sea_breeze_diag.f90:375
  subroutine get_dist(coast, landfrac, lon, lat, maxdist, cdist, halo_size)
    real, intent(in), contiguous :: coast(:, :), landfrac(:, :), lon(:), lat(:)
    real, intent(in) :: maxdist
    real, intent(out), contiguous :: cdist(:, :)
    integer, intent(in) :: halo_size
    call check(sbd_get_dist_f32(coast, landfrac, lon, lat, maxdist, cdist, size(lon, 1), size(lat, 1), &
               halo_size), 'get_dist')
  end subroutine get_dist

  ! generic/sea_breeze_diag.f90:457
  subroutine sigmoid(ary, sm)
    real, intent(in), contiguous :: ary(:, :)
    real, intent(out), contiguous :: sm(:, :)
    call check(sbd_sigmoid_f32(ary, sm, size(ary, 1), size(ary, 2), c_null_ptr), 'sigmoid')
  end subroutine sigmoid

end module sea_breeze_diag_mod
