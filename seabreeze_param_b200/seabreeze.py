"""Drop-in for the f2py extension module ``seabreeze`` (python_wrapper/setup.py:12-16) built from
python_wrapper/seabreezediag/{seabreeze_diag_python,sobel}.f90.

Same callables, positional order, keyword names, defaults, dtypes (float32) and Fortran-ordered
outputs as the f2py signatures:

    output = diag(timestep_number,p,z,std,theta,v,u,cdist,windspeed,winddir,thc,
                  [target_plev=700.,thresh_wind=11,thresh_winddir=90.,thresh_windch=5.,
                   thresh_thc=0.75,target_time=6,maxdist=180,timestep=24,nps,nlons,nlats])
    coast  = get_edges(lsm,ci,[nlons,nlats])
    cdist  = get_dist(coast,mask,lon,lat,[nlons,nlats,maxdist=180])
    sm     = sigmoid(ary,[nlons,nlats])
    nt     = get_threads()

All arithmetic runs on the GPU through include/sbd.h (sbd_py_*).
"""
import ctypes as C

import numpy as np

from . import _capi


def _f32(a, ndim, name):
    """f2py's input conversion: float32, Fortran-contiguous, rank checked."""
    arr = np.asarray(a)
    if arr.ndim != ndim:
        raise ValueError("seabreeze: %s must have rank %d (got %d)" % (name, ndim, arr.ndim))
    return np.asfortranarray(arr, dtype=np.float32)


def _chk(arr, shape, name):
    if tuple(arr.shape) != tuple(shape):
        raise ValueError("seabreeze: shape of %s is %s, expected %s" % (name, arr.shape, tuple(shape)))


def get_threads():
    """nt = get_threads()  (sobel.f90:195-206; here: number of CUDA devices)."""
    return int(_capi.load().sbd_get_threads())


def get_edges(lsm, ci, nlons=None, nlats=None):
    """coast = get_edges(lsm,ci,[nlons,nlats])  (sobel.f90:19-89)."""
    lsm = _f32(lsm, 2, "lsm")
    nlons = lsm.shape[0] if nlons is None else nlons
    nlats = lsm.shape[1] if nlats is None else nlats
    _chk(lsm, (nlons, nlats), "lsm")
    ci = _f32(ci, 2, "ci")
    _chk(ci, (nlons, nlats), "ci")
    coast = np.empty((nlons, nlats), dtype=np.float32, order="F")
    lib = _capi.load()
    _capi.check(lib.sbd_py_get_edges_f32(_capi.ptr(lsm), _capi.ptr(ci), nlons, nlats, _capi.ptr(coast)),
                "get_edges")
    return coast


def get_dist(coast, mask, lon, lat, nlons=None, nlats=None, maxdist=180.0):
    """cdist = get_dist(coast,mask,lon,lat,[nlons,nlats,maxdist])  (sobel.f90:91-193)."""
    coast = _f32(coast, 2, "coast")
    nlons = coast.shape[0] if nlons is None else nlons
    nlats = coast.shape[1] if nlats is None else nlats
    _chk(coast, (nlons, nlats), "coast")
    mask = _f32(mask, 2, "mask")
    _chk(mask, (nlons, nlats), "mask")
    lon = _f32(lon, 1, "lon")
    lat = _f32(lat, 1, "lat")
    _chk(lon, (nlons,), "lon")
    _chk(lat, (nlats,), "lat")
    cdist = np.empty((nlons, nlats), dtype=np.float32, order="F")
    lib = _capi.load()
    k = C.c_int(0)
    _capi.check(lib.sbd_py_get_dist_f32(_capi.ptr(coast), _capi.ptr(mask), _capi.ptr(lon), _capi.ptr(lat),
                                        nlons, nlats, float(maxdist), _capi.ptr(cdist), C.byref(k)),
                "get_dist")
    return cdist


def sigmoid(ary, nlons=None, nlats=None):
    """sm = sigmoid(ary,[nlons,nlats])  (seabreeze_diag_python.f90:287-311)."""
    ary = _f32(ary, 2, "ary")
    nlons = ary.shape[0] if nlons is None else nlons
    nlats = ary.shape[1] if nlats is None else nlats
    _chk(ary, (nlons, nlats), "ary")
    sm = np.empty((nlons, nlats), dtype=np.float32, order="F")
    lib = _capi.load()
    _capi.check(lib.sbd_py_sigmoid_f32(_capi.ptr(ary), nlons, nlats, _capi.ptr(sm)), "sigmoid")
    return sm


def diag(timestep_number, p, z, std, theta, v, u, cdist, windspeed, winddir, thc, target_plev=700.0,
         thresh_wind=11.0, thresh_winddir=90.0, thresh_windch=5.0, thresh_thc=0.75, target_time=6.0,
         maxdist=180.0, timestep=24.0, nps=None, nlons=None, nlats=None, flags=None):
    """output = diag(...)  (seabreeze_diag_python.f90:49-285).

    output is float32, Fortran-ordered (nlons,nlats,4): sb_con, t0, windspeed, winddir.
    ``flags`` (extension, optional uint8 (nlons,nlats) F-ordered array): bit0 coastal zone,
    bit1 trigger."""
    p = _f32(p, 1, "p")
    nps = p.shape[0] if nps is None else nps
    z = _f32(z, 2, "z")
    nlons = z.shape[0] if nlons is None else nlons
    nlats = z.shape[1] if nlats is None else nlats
    _chk(p, (nps,), "p")
    arrs2 = {}
    for name, a in (("z", z), ("std", std), ("theta", theta), ("cdist", cdist),
                    ("windspeed", windspeed), ("winddir", winddir), ("thc", thc)):
        a = _f32(a, 2, name)
        _chk(a, (nlons, nlats), name)
        arrs2[name] = a
    v = _f32(v, 3, "v")
    u = _f32(u, 3, "u")
    _chk(v, (nlons, nlats, nps), "v")
    _chk(u, (nlons, nlats, nps), "u")
    out = np.empty((nlons, nlats, 4), dtype=np.float32, order="F")
    fl = None
    if flags is not None:
        if flags.dtype != np.uint8 or flags.shape != (nlons, nlats) or not flags.flags.f_contiguous:
            raise ValueError("flags must be a Fortran-ordered uint8 (nlons,nlats) array")
        fl = flags
    lib = _capi.load()
    q = _capi.ptr
    _capi.check(lib.sbd_py_diag_f32(int(timestep_number), q(p), q(arrs2["z"]), q(arrs2["std"]),
                                    q(arrs2["theta"]), q(v), q(u), q(arrs2["cdist"]),
                                    q(arrs2["windspeed"]), q(arrs2["winddir"]), q(arrs2["thc"]),
                                    float(target_plev), float(thresh_wind), float(thresh_winddir),
                                    float(thresh_windch), float(thresh_thc), float(target_time),
                                    float(maxdist), float(timestep), nps, nlons, nlats, q(out), q(fl)),
                "diag")
    return out
