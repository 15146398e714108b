"""ctypes binding of the CPU oracle (oracle/sbd_oracle.cpp).  TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the
product package.  Arrays are float32, C-ordered (ny, nx[...]) == Fortran (nx, ny[...])."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libsbd_oracle.so")
_lib = None


def build(force=False):
    src = [os.path.join(HERE, "sbd_oracle.cpp"), os.path.join(HERE, "..", "include", "sbd_math.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(s) <= os.path.getmtime(LIB) for s in src):
        return LIB
    res = subprocess.run(["make", "-C", HERE] + (["-B"] if force else []), capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        fp, i, f, vp = C.c_void_p, C.c_int, C.c_float, C.c_void_p
        L.orc_set_math.argtypes = [i]
        L.orc_set_threads.argtypes = [i]
        L.orc_sigmoid.argtypes = [fp, i, i, fp, fp]
        L.orc_sigmoid_strided.argtypes = [fp, i, i, i, fp, i, fp]
        L.orc_py_get_edges.argtypes = [fp, fp, i, i, fp]
        L.orc_py_dist_radius.argtypes = [fp, fp, i, i, f]
        L.orc_py_get_dist.argtypes = [fp, fp, fp, fp, i, i, f, i, i, fp]
        L.orc_py_diag.argtypes = [i] + [fp] * 10 + [f] * 8 + [i, i, i, i, fp, vp]
        L.orc_swap_bounds.argtypes = [fp, i, i, i]
        L.orc_get_edges.argtypes = [fp, fp, fp, i, i, i, i, vp]
        L.orc_get_dist.argtypes = [fp, fp, fp, fp, f, fp, i, i, i, i]
        L.orc_sobel.argtypes = [fp, i, i, i, fp, fp]
        L.orc_window_info.argtypes = [fp, i, i, i, f, vp, vp, vp]
        L.orc_seabreeze_diag.argtypes = [f, i] + [fp] * 11 + [i, i, i, i, i, i, vp]
        L.orc_um_swap_bounds.argtypes = [fp, i, i, i, i]
        L.orc_um_get_dist.argtypes = [fp, fp, fp, fp, i, i, i, i, f]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data


def _c(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def set_math(mode):
    """0 = platform libm, 1 = pinned (include/sbd_math.h)."""
    lib().orc_set_math(int(mode))


def set_threads(n):
    lib().orc_set_threads(int(n))


def max_threads():
    return int(lib().orc_max_threads())


def sigmoid(ary):
    ary = _c(ary)
    ny, nx = ary.shape
    sm = np.empty_like(ary)
    st = np.zeros(4, dtype=np.float32)
    lib().orc_sigmoid(_p(ary), nx, ny, _p(sm), _p(st))
    return sm, st


def py_get_edges(lsm, ci):
    lsm, ci = _c(lsm), _c(ci)
    ny, nx = lsm.shape
    coast = np.empty_like(lsm)
    lib().orc_py_get_edges(_p(lsm), _p(ci), nx, ny, _p(coast))
    return coast


def py_dist_radius(lon, lat, maxdist=180.0):
    lon, lat = _c(lon), _c(lat)
    return int(lib().orc_py_dist_radius(_p(lon), _p(lat), len(lon), len(lat), float(maxdist)))


def py_get_dist(coast, mask, lon, lat, maxdist=180.0, k=-1, literal=True):
    coast, mask, lon, lat = _c(coast), _c(mask), _c(lon), _c(lat)
    ny, nx = coast.shape
    cd = np.empty_like(coast)
    rc = lib().orc_py_get_dist(_p(coast), _p(mask), _p(lon), _p(lat), nx, ny, float(maxdist), int(k),
                               1 if literal else 0, _p(cd))
    if rc:
        raise RuntimeError("orc_py_get_dist status %d" % rc)
    return cd


def py_diag(step, p, z, std, theta, v, u, cdist, ws, wd, thc=None, target_plev=700.0, thresh_wind=11.0,
            thresh_winddir=90.0, thresh_windch=5.0, thresh_thc=0.75, target_time=6.0, maxdist=180.0,
            timestep=24.0, literal=True, want_flags=True):
    """Returns output (4, ny, nx) and flags (ny, nx)."""
    p, z, std, theta, v, u, cdist, ws, wd = map(_c, (p, z, std, theta, v, u, cdist, ws, wd))
    thc = _c(thc) if thc is not None else np.zeros_like(z)
    ny, nx = z.shape
    nps = p.shape[0]
    out = np.empty((4, ny, nx), dtype=np.float32)
    fl = np.zeros((ny, nx), dtype=np.uint8) if want_flags else None
    rc = lib().orc_py_diag(int(step), _p(p), _p(z), _p(std), _p(theta), _p(v), _p(u), _p(cdist), _p(ws), _p(wd),
                           _p(thc), target_plev, thresh_wind, thresh_winddir, thresh_windch, thresh_thc,
                           target_time, maxdist, timestep, nps, nx, ny, 1 if literal else 0, _p(out), _p(fl))
    if rc:
        raise RuntimeError("orc_py_diag status %d" % rc)
    return out, fl


def swap_bounds(field, h):
    assert field.dtype == np.float32 and field.flags.c_contiguous
    ny, nx = field.shape[0] - 2 * h, field.shape[1] - 2 * h
    lib().orc_swap_bounds(_p(field), nx, ny, h)
    return field


def get_edges(icefrac, landfrac, h, swap=True, mask_in=None):
    """Returns (mask halo'd, m1)."""
    ice, land = _c(icefrac), _c(landfrac)
    ny, nx = land.shape
    mask = np.zeros((ny + 2 * h, nx + 2 * h), dtype=np.float32) if mask_in is None else _c(mask_in).copy()
    m1 = np.zeros((ny, nx), dtype=np.uint8)
    rc = lib().orc_get_edges(_p(mask), _p(ice), _p(land), nx, ny, h, 1 if swap else 0, _p(m1))
    if rc:
        raise RuntimeError("orc_get_edges status %d" % rc)
    return mask, m1


def sobel(mask_h, h):
    """(px, py) of generic/sea_breeze_diag.f90:350-357 on a halo'd binary mask."""
    m = _c(mask_h)
    ny, nx = m.shape[0] - 2 * h, m.shape[1] - 2 * h
    px = np.zeros((ny, nx), np.float32); py = np.zeros((ny, nx), np.float32)
    rc = lib().orc_sobel(_p(m), nx, ny, h, _p(px), _p(py))
    if rc:
        raise RuntimeError("orc_sobel status %d" % rc)
    return px, py


def window_info(mask_h, h, maxdist=180.0):
    """(nn, n_land, n_sea) at the last pass of the growing-window loop, G:188-214; int32 (ny, nx)."""
    m = _c(mask_h)
    ny, nx = m.shape[0] - 2 * h, m.shape[1] - 2 * h
    out = [np.zeros((ny, nx), np.int32) for _ in range(3)]
    rc = lib().orc_window_info(_p(m), nx, ny, h, float(maxdist), *[o.ctypes.data for o in out])
    if rc:
        raise RuntimeError("orc_window_info status %d" % rc)
    return out


def get_dist(coast_h, landfrac, lon, lat, maxdist, h, literal=True):
    coast_h, land, lon, lat = _c(coast_h), _c(landfrac), _c(lon), _c(lat)
    ny, nx = land.shape
    cd = np.empty((ny, nx), dtype=np.float32)
    rc = lib().orc_get_dist(_p(coast_h), _p(land), _p(lon), _p(lat), float(maxdist), _p(cd), nx, ny, h,
                            1 if literal else 0)
    if rc:
        raise RuntimeError("orc_get_dist status %d" % rc)
    return cd


def seabreeze_diag(timestep, step, p, u, v, theta, mask, z, sigma, ws, wd, thc, sb, h, literal=True,
                   level_scan=0):
    """In-place on ws, wd, thc, sb (float32 C-contiguous (ny,nx)); returns flags."""
    p, u, v, theta, mask, z, sigma = map(_c, (p, u, v, theta, mask, z, sigma))
    nz, ny, nx = p.shape
    for a in (ws, wd, thc, sb):
        assert a.dtype == np.float32 and a.flags.c_contiguous and a.shape == (ny, nx)
    fl = np.zeros((ny, nx), dtype=np.uint8)
    rc = lib().orc_seabreeze_diag(float(timestep), int(step), _p(p), _p(u), _p(v), _p(theta), _p(mask), _p(z),
                                  _p(sigma), _p(ws), _p(wd), _p(thc), _p(sb), nx, ny, nz, h,
                                  1 if literal else 0, int(level_scan), _p(fl))
    if rc:
        raise RuntimeError("orc_seabreeze_diag status %d" % rc)
    return fl


def um_swap_bounds(field, nx, ny, hi, hj):
    assert field.dtype == np.float32 and field.flags.c_contiguous and field.shape == (ny + 2 * hj, nx + 2 * hi)
    lib().orc_um_swap_bounds(_p(field), nx, ny, hi, hj)
    return field


def um_get_dist(landfrac, coast_l, lat2d, lon2d, hi, hj, maxdist=180.0):
    """UM get_dist(landfrac, coast) (UM/vn10.7/sea_breeze_diag.F90:448-601), literal scatter; returns the new halo'd
    coast array (signed distance)."""
    land, lat2d, lon2d = _c(landfrac), _c(lat2d), _c(lon2d)
    ny, nx = land.shape
    out = _c(coast_l).copy()
    assert out.shape == (ny + 2 * hj, nx + 2 * hi)
    rc = lib().orc_um_get_dist(_p(land), _p(out), _p(lat2d), _p(lon2d), nx, ny, hi, hj, float(maxdist))
    if rc:
        raise RuntimeError("orc_um_get_dist status %d" % rc)
    return out
