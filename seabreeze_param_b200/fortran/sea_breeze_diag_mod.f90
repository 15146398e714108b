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
    ! The C side reads (nx + 2h) x (ny + 2h) elements of every neighbour-read field: refuse anything else.  NOTE
    ! that generic/get_all_fields_mod.f90:17-19 declares mask, theta as (nx+halo, ny+halo) and z, sigma as (nx, ny):
    ! those arrays are NOT conformable with this interface (nor with G:167 itself, SURVEY.md A.6) -- allocate all
    ! four with 2*halo_size extra points in both directions, as the UM does (U:78-117).
    if (h < 1 .or. size(mask, 1) /= nx + 2 * h .or. size(mask, 2) /= ny + 2 * h) &
      error stop 'sea_breeze_diag_mod: mask must be (nx + 2*halo, ny + 2*halo), halo >= 1'
    if (any(shape(theta) /= shape(mask)) .or. any(shape(z) /= shape(mask)) .or. any(shape(sigma) /= shape(mask))) &
      error stop 'sea_breeze_diag_mod: theta, z, sigma and mask must share the shape (nx + 2*halo, ny + 2*halo)'
    if (any(shape(u) /= shape(p)) .or. any(shape(v) /= shape(p))) &
      error stop 'sea_breeze_diag_mod: p, u, v must share one shape'
    if (any(shape(windspeed) /= (/nx, ny/)) .or. any(shape(winddir) /= (/nx, ny/)) .or. &
        any(shape(thc) /= (/nx, ny/)) .or. any(shape(sb_con) /= (/nx, ny/))) &
      error stop 'sea_breeze_diag_mod: windspeed, winddir, thc, sb_con must be (nx, ny)'
    call check(sbd_seabreeze_diag_f32(timestep, timestep_number, p, u, v, theta, mask, z, sigma, &
               windspeed, winddir, thc, sb_con, nx, ny, nz, h, c_null_ptr, c_null_ptr), 'seabreeze_diag')
  end subroutine seabreeze_diag

  ! generic/sea_breeze_diag.f90:273
  subroutine get_edges(mask, icefrac, landfrac, halo_size)
    real, intent(inout), contiguous :: mask(:, :)
    real, intent(in), contiguous :: icefrac(:, :), landfrac(:, :)
    integer, intent(in) :: halo_size
    if (halo_size < 1 .or. size(mask, 1) /= size(landfrac, 1) + 2 * halo_size .or. &
        size(mask, 2) /= size(landfrac, 2) + 2 * halo_size) &
      error stop 'sea_breeze_diag_mod: get_edges needs mask(nx + 2*halo_size, ny + 2*halo_size)'
    if (any(shape(icefrac) /= shape(landfrac))) error stop 'sea_breeze_diag_mod: icefrac and landfrac differ in shape'
    call check(sbd_get_edges_f32(mask, icefrac, landfrac, size(landfrac, 1), size(landfrac, 2), &
               halo_size, 1_c_int, c_null_ptr), 'get_edges')
  end subroutine get_edges

  ! generic/sea_breeze_diag.f90:375
  subroutine get_dist(coast, landfrac, lon, lat, maxdist, cdist, halo_size)
    real, intent(in), contiguous :: coast(:, :), landfrac(:, :), lon(:), lat(:)
    real, intent(in) :: maxdist
    real, intent(out), contiguous :: cdist(:, :)
    integer, intent(in) :: halo_size
    if (size(coast, 1) /= size(lon, 1) + 2 * halo_size .or. size(coast, 2) /= size(lat, 1) + 2 * halo_size) &
      error stop 'sea_breeze_diag_mod: get_dist needs coast(nx + 2*halo_size, ny + 2*halo_size) (output of get_edges)'
    if (size(landfrac, 1) /= size(lon, 1) .or. size(landfrac, 2) /= size(lat, 1) .or. &
        any(shape(cdist) /= shape(landfrac))) &
      error stop 'sea_breeze_diag_mod: get_dist needs landfrac(nx, ny) and cdist(nx, ny)'
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
