"""Host-side mirror of the UM hook's module (UM/vn10.7/sea_breeze_diag.F90, "U") for Python callers and tests:
same routine names and argument order as U, arrays as numpy arrays (host) or torch CUDA tensors whose memory is
the Fortran array's (C-ordered (rows + 2 halo_j, row_length + 2 halo_i) == Fortran (i, j) with the UM's bounds).
The Fortran side is fortran/sea_breeze_diag_um.F90; both call include/sbd.h's sbd_um_* entry points."""
import ctypes as C

import numpy as np

from . import _capi


class UmError(RuntimeError):
    """`error /= 0` of the UM routine (U:102)"""


def seabreeze_diag(timestep, timestep_number, p, u, v, theta, z, sigma, mask, windspeed, winddir, thc, sb_con,
                   halo_s, halo_l, theta_to_t0=True, params=None, flags=None):
    """call seabreeze_diag(timestep, timestep_number, p, u, v, theta, z, sigma, mask, windspeed, winddir, thc,
    sb_con, error)  (U:55-326).  Returns `error` (0 = ok, 1 = bad dimensions, U:198-201); other failures raise."""
    nz, ny, nx = p.shape
    for name, a, hh in (("theta", theta, halo_s), ("z", z, halo_s), ("sigma", sigma, halo_s), ("mask", mask, halo_l)):
        if tuple(a.shape) != (ny + 2 * hh, nx + 2 * hh):
            raise ValueError("%s has shape %s, expected %s" % (name, tuple(a.shape), (ny + 2 * hh, nx + 2 * hh)))
    lib = _capi.load()
    q = _capi.ptr
    prm = C.byref(params) if params is not None else None
    rc = lib.sbd_um_seabreeze_diag_f32(float(timestep), int(timestep_number), q(p), q(u), q(v), q(theta), q(z), q(sigma),
                                       q(mask), q(windspeed), q(winddir), q(thc), q(sb_con), nx, ny, nz, int(halo_s),
                                       int(halo_l), 1 if theta_to_t0 else 0, prm, q(flags))
    if rc not in (0, 1):
        _capi.check(rc, "um seabreeze_diag")
    return rc


def get_edges(mask, icefrac, landfrac, halo_i, halo_j):
    """call get_edges(mask, icefrac, landfrac)  (U:328-446); mask: tdims_l plane, written in place."""
    ny, nx = landfrac.shape
    assert tuple(mask.shape) == (ny + 2 * halo_j, nx + 2 * halo_i)
    lib = _capi.load()
    _capi.check(lib.sbd_um_get_edges_f32(_capi.ptr(mask), _capi.ptr(icefrac), _capi.ptr(landfrac), nx, ny, halo_i, halo_j),
                "um get_edges")
    return mask


def get_dist(landfrac, coast, true_latitude, true_longitude, halo_i, halo_j, maxdist=180.0):
    """call get_dist(landfrac, coast)  (U:448-601); landfrac, coast: tdims_l planes, coast in place."""
    ny, nx = true_latitude.shape
    assert tuple(coast.shape) == (ny + 2 * halo_j, nx + 2 * halo_i) == tuple(landfrac.shape)
    lib = _capi.load()
    _capi.check(lib.sbd_um_get_dist_f32(_capi.ptr(landfrac), _capi.ptr(coast), _capi.ptr(true_latitude),
                                        _capi.ptr(true_longitude), nx, ny, halo_i, halo_j, float(maxdist)), "um get_dist")
    return coast


def swap_bounds(field, row_len, rows, levels, halo_i, halo_j, fld_type=None, l_vector=False):
    """call swap_bounds(field, row_len, rows, levels, halo_i, halo_j, fld_type, l_vector)  (U:408-410)"""
    lib = _capi.load()
    _capi.check(lib.sbd_um_swap_bounds_f32(_capi.ptr(field), row_len, rows, levels, halo_i, halo_j), "um swap_bounds")
    return field
