!------------------------------------------------------------------------------
! halo_exchange_mod -- swap_bounds(field, halo_size) over the CUDA C-ABI.
! Replaces the empty stub generic/halo_exchange_mod.f90:12-17 for a single
! domain: east-west periodic, north-south edge-replicated halo fill
! (include/sbd.h: sbd_swap_bounds_f32).  Multi-GPU row blocks exchange their
! north-south halos with NCCL send/recv (seabreeze_param_b200/halo.py).
!------------------------------------------------------------------------------
module halo_exchange_mod
  use, intrinsic :: iso_c_binding, only : c_int, c_float
  implicit none
  interface
    function sbd_swap_bounds_f32(field, nx, ny, h) bind(C, name="sbd_swap_bounds_f32") result(rc)
      import :: c_int, c_float
      real(c_float), intent(inout) :: field(*)
      integer(c_int), value :: nx, ny, h
      integer(c_int) :: rc
    end function sbd_swap_bounds_f32
  end interface
contains
  subroutine swap_bounds(field, halo_size)
    integer, intent(in) :: halo_size
    real, intent(inout), contiguous :: field(:, :)
    integer(c_int) :: rc
    rc = sbd_swap_bounds_f32(field, size(field, 1) - 2 * halo_size, size(field, 2) - 2 * halo_size, halo_size)
    if (rc /= 0) error stop 1
  end subroutine swap_bounds
end module halo_exchange_mod
