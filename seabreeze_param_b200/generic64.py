"""64-bit REAL mirrors of sea_breeze_diag_mod for hosts built with promoted reals (UM -r8;
UM/vn10.7/sea_breeze_diag.F90:64 real(kind=jprb)).  Same routines and argument order as
seabreeze_param_b200.generic, float64 numpy arrays, through the _f64 entry points of the C-ABI (include/sbd.h):
sigmoid and seabreeze_diag run in binary64 throughout (csrc/sbd_f64.cuh); get_edges, get_dist and swap_bounds --
static, one-off work on small integers / coast distances -- are narrowed to fp32 on the device and widened exactly."""
import ctypes as C

import numpy as np

from . import _capi


def _chk(a, shape, name):
    if tuple(a.shape) != tuple(shape) or a.dtype != np.float64 or not a.flags.c_contiguous:
        raise ValueError("%s must be a C-contiguous float64 array of shape %s" % (name, tuple(shape)))
    return a.ctypes.data


def get_edges(mask, icefrac, landfrac, halo_size, swap=True):
    ny, nx = landfrac.shape
    h = int(halo_size)
    lib = _capi.load()
    _capi.check(lib.sbd_get_edges_f64(_chk(mask, (ny + 2 * h, nx + 2 * h), "mask"), _chk(icefrac, (ny, nx), "icefrac"),
                                      _chk(landfrac, (ny, nx), "landfrac"), nx, ny, h, 1 if swap else 0), "get_edges_f64")
    return mask


def get_dist(coast, landfrac, lon, lat, maxdist, cdist, halo_size):
    ny, nx = landfrac.shape
    h = int(halo_size)
    lib = _capi.load()
    _capi.check(lib.sbd_get_dist_f64(_chk(coast, (ny + 2 * h, nx + 2 * h), "coast"), _chk(landfrac, (ny, nx), "landfrac"),
                                     _chk(lon, (nx,), "lon"), _chk(lat, (ny,), "lat"), float(maxdist),
                                     _chk(cdist, (ny, nx), "cdist"), nx, ny, h), "get_dist_f64")
    return cdist


def sigmoid(ary, sm=None):
    ny, nx = ary.shape
    if sm is None:
        sm = np.empty_like(ary)
    lib = _capi.load()
    _capi.check(lib.sbd_sigmoid_f64(_chk(ary, (ny, nx), "ary"), _chk(sm, (ny, nx), "sm"), nx, ny), "sigmoid_f64")
    return sm


def swap_bounds(field, halo_size):
    h = int(halo_size)
    ny, nx = field.shape[0] - 2 * h, field.shape[1] - 2 * h
    lib = _capi.load()
    _capi.check(lib.sbd_swap_bounds_f64(_chk(field, field.shape, "field"), nx, ny, h), "swap_bounds_f64")
    return field


def seabreeze_diag(timestep, timestep_number, p, u, v, theta, mask, z, sigma, windspeed, winddir, thc, sb_con,
                   params=None, flags=None):
    nz, ny, nx = p.shape
    h = (theta.shape[1] - nx) // 2
    q3 = [_chk(a, (nz, ny, nx), n) for n, a in (("p", p), ("u", u), ("v", v))]
    qh = [_chk(a, (ny + 2 * h, nx + 2 * h), n) for n, a in (("theta", theta), ("mask", mask), ("z", z), ("sigma", sigma))]
    q2 = [_chk(a, (ny, nx), n) for n, a in (("windspeed", windspeed), ("winddir", winddir), ("thc", thc), ("sb_con", sb_con))]
    lib = _capi.load()
    prm = C.byref(params) if params is not None else None
    _capi.check(lib.sbd_seabreeze_diag_f64(float(timestep), int(timestep_number), *q3, *qh, *q2, nx, ny, nz, h, prm,
                                           _capi.ptr(flags)), "seabreeze_diag_f64")
    return sb_con
