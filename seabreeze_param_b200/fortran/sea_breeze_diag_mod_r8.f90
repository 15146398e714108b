!------------------------------------------------------------------------------
! sea_breeze_diag_mod (64-bit REAL build) -- the same shim as sea_breeze_diag_mod.f90
! for hosts compiled with promoted reals (gfortran -fdefault-real-8, ifort -r8,
! the UM's real(kind=jprb), UM/vn10.7/sea_breeze_diag.F90:64): the dummies are
! real(c_double) and the calls go to the _f64 twins of the C-ABI (include/sbd.h),
! which narrow to fp32 on the device, run the fp32 kernels and widen the results.
! Use this file INSTEAD of sea_breeze_diag_mod.f90 in such a build.  For the UM
! hook, `error` is the status the C functions return (0 = ok; 1 = bad dimensions,
! as UM/vn10.7/sea_breeze_diag.F90:198-201).
!
! NOTE: no Fortran compiler exists in the build image of this repository, so this
! file is compiled by the integrator; tests/test_gpu_f64.py exercises the same
! entry points through ctypes.
!------------------------------------------------------------------------------
module sea_breeze_diag_mod
  use, intrinsic :: iso_c_binding, only : c_int, c_double, c_ptr, c_null_ptr
  implicit none
  private
  public :: seabreeze_diag, get_edges, get_dist, sigmoid

  interface
    function sbd_get_edges_f64(mask, icefrac, landfrac, nx, ny, h, swap) &
        bind(C, name="sbd_get_edges_f64") result(rc)
      import :: c_int, c_double
      real(c_double), intent(inout) :: mask(*)
      real(c_double), intent(in) :: icefrac(*), landfrac(*)
      integer(c_int), value :: nx, ny, h, swap
      integer(c_int) :: rc
    end function sbd_get_edges_f64

    function sbd_get_dist_f64(coast, landfrac, lon, lat, maxdist, cdist, nx, ny, h) &
        bind(C, name="sbd_get_dist_f64") result(rc)
      import :: c_int, c_double
      real(c_double), intent(in) :: coast(*), landfrac(*), lon(*), lat(*)
      real(c_double), value :: maxdist
      real(c_double), intent(out) :: cdist(*)
      integer(c_int), value :: nx, ny, h
      integer(c_int) :: rc
    end function sbd_get_dist_f64

    function sbd_sigmoid_f64(ary, sm, nx, ny) bind(C, name="sbd_sigmoid_f64") result(rc)
      import :: c_int, c_double
      real(c_double), intent(in) :: ary(*)
      real(c_double), intent(out) :: sm(*)
      integer(c_int), value :: nx, ny
      integer(c_int) :: rc
    end function sbd_sigmoid_f64

    function sbd_seabreeze_diag_f64(timestep, timestep_number, p, u, v, theta, mask, z, sigma, &
        windspeed, winddir, thc, sb_con, nx, ny, nz, h, prm, flags) &
        bind(C, name="sbd_seabreeze_diag_f64") result(rc)
      import :: c_int, c_double, c_ptr
      real(c_double), value :: timestep
      integer(c_int), value :: timestep_number
      real(c_double), intent(in) :: p(*), u(*), v(*), theta(*), mask(*), z(*), sigma(*)
      real(c_double), intent(inout) :: windspeed(*), winddir(*), thc(*), sb_con(*)
      integer(c_int), value :: nx, ny, nz, h
      type(c_ptr), value :: prm, flags
      integer(c_int) :: rc
    end function sbd_seabreeze_diag_f64
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
    real(c_double), intent(in) :: timestep
    integer, intent(in) :: timestep_number
    real(c_double), intent(in), contiguous :: p(:, :, :), u(:, :, :), v(:, :, :)
    real(c_double), intent(in), contiguous :: theta(:, :), z(:, :), sigma(:, :), mask(:, :)
    real(c_double), intent(inout), contiguous :: windspeed(:, :), winddir(:, :), thc(:, :), sb_con(:, :)
    integer(c_int) :: nx, ny, nz, h
    nx = size(p, 1); ny = size(p, 2); nz = size(p, 3)
    h = (size(mask, 1) - nx) / 2          ! halo'd neighbour fields (UM tdims_s / tdims_l)
    call check(sbd_seabreeze_diag_f64(timestep, timestep_number, p, u, v, theta, mask, z, sigma, &
               windspeed, winddir, thc, sb_con, nx, ny, nz, h, c_null_ptr, c_null_ptr), 'seabreeze_diag')
  end subroutine seabreeze_diag

  ! generic/sea_breeze_diag.f90:273
  subroutine get_edges(mask, icefrac, landfrac, halo_size)
    real(c_double), intent(inout), contiguous :: mask(:, :)
    real(c_double), intent(in), contiguous :: icefrac(:, :), landfrac(:, :)
    integer, intent(in) :: halo_size
    call check(sbd_get_edges_f64(mask, icefrac, landfrac, size(landfrac, 1), size(landfrac, 2), &
               halo_size, 1_c_int), 'get_edges')
  end subroutine get_edges

  ! generic/sea_breeze_diag.f90:375
  subroutine get_dist(coast, landfrac, lon, lat, maxdist, cdist, halo_size)
    real(c_double), intent(in), contiguous :: coast(:, :), landfrac(:, :), lon(:), lat(:)
    real(c_double), intent(in) :: maxdist
    real(c_double), intent(out), contiguous :: cdist(:, :)
    integer, intent(in) :: halo_size
    call check(sbd_get_dist_f64(coast, landfrac, lon, lat, maxdist, cdist, size(lon, 1), size(lat, 1), &
               halo_size), 'get_dist')
  end subroutine get_dist

  ! generic/sea_breeze_diag.f90:457
  subroutine sigmoid(ary, sm)
    real(c_double), intent(in), contiguous :: ary(:, :)
    real(c_double), intent(out), contiguous :: sm(:, :)
    call check(sbd_sigmoid_f64(ary, sm, size(ary, 1), size(ary, 2)), 'sigmoid')
  end subroutine sigmoid

end module sea_breeze_diag_mod
