"""Python API of the reference (python_wrapper/seabreezediag/__init__.py, "I") on the CUDA C-ABI.

``diag`` has the reference's call signature and return tuple (I:91, I:263) and the same per-record
sequence get_edges -> get_dist -> diag (I:228-239) with the state threaded from record to record
(I:244).  ``c2f`` is the reference's order-flipping view (I:13-40).

Behaviour kept from the reference:
  * inputs are C-ordered (lat, lon) / (pres, lat, lon); ``c2f`` turns them into the (lon, lat[, pres])
    Fortran views the kernels consume, without copying;
  * the "thermal heating contrast" returned / carried is plane 1 of the kernel output, i.e. t0
    (I:244 assigns out[1]); the kernel never reads it;
  * missing state after the first record warns (I:204-215).
Defined where the reference is broken: ``ci=None`` means "no sea ice" (I:216-218 raises NameError),
and inputs without a time axis are treated as one record (I:246-262 references an undefined name).

The reference file itself also runs unchanged: put ``seabreeze_param_b200/dropin`` (a module named
``seabreeze``) on PYTHONPATH -- see INTEGRATION.md.
"""
import warnings

import numpy as np

from . import _capi
from . import seabreeze as _sb


def c2f(array):
    """Reinterpret a C-ordered array as the Fortran-ordered array with reversed shape (I:13-40)."""
    return array.ravel().reshape(array.shape[::-1], order="F")


def _pop(kw, key):
    return kw.pop(key, None)


def _state(value, like, label, tt):
    if value is not None:
        return value
    if tt > 1:
        warnings.warn("%s should be given from previous timestep" % label)
    return np.zeros_like(like)


def _unmask(a):
    return a.filled(0) if hasattr(a, "filled") else a


def coast_distance(lsm, ice, lon, lat):
    """Signed distance to the coast, C-ordered (lat, lon): get_dist(get_edges(lsm, ice), lsm) with
    get_dist's default maxdist, exactly as I:228 calls it (the maxdist keyword only reaches diag)."""
    lsm_f = c2f(np.ascontiguousarray(lsm))
    ice_f = c2f(np.ascontiguousarray(ice))
    coast = _sb.get_edges(lsm_f, ice_f)
    return c2f(_sb.get_dist(coast, lsm_f, lon, lat))


def diag(tt, lsm, z, std, lon, lat, pres, *args, **kwargs):
    """tt, sb_con, thc, ws, wd = diag(tt, lsm, z, std, lon, lat, pres, u, v, t, ci,
                                      ws=, wd=, thc=, meta=, **tunables)

    u, v: ([time], pres, lat, lon); t, ci: ([time], lat, lon); tunables are the keyword arguments
    of the kernel (target_plev, thresh_wind, thresh_winddir, thresh_windch, thresh_thc,
    target_time, maxdist, timestep).  Returns the advanced timestep counter, sb_con with a leading
    time axis, and the state (t0, windspeed, wind direction) after the last record."""
    ws, wd, thc, meta = (_pop(kwargs, k) for k in ("ws", "wd", "thc", "meta"))
    if meta is None:
        u, v, t, ci = args
    else:
        u, v, t, ci = meta.u, meta.v, meta.theta, getattr(meta, "ci", None)
    tt = max(1, tt)
    ws = _state(ws, lsm, "Windspeed", tt)
    wd = _state(wd, lsm, "Wind direction", tt)
    thc = _state(thc, lsm, "Heating contrast", tt)

    timed = len(v.shape) > 3
    nrec = len(v) if timed else 1
    nlat, nlon = t.shape[-2], t.shape[-1]
    result = np.zeros([4, nrec, nlat, nlon])
    pres = np.ascontiguousarray(np.asarray(pres))
    static_dist = None
    if ci is None:
        static_dist = coast_distance(lsm, np.zeros_like(lsm), lon, lat)
    for rec in range(nrec):
        pick = (lambda a: a[rec]) if timed else (lambda a: a[:])
        dist = static_dist
        if dist is None:
            dist = coast_distance(lsm, _unmask(pick(ci)), lon, lat)
        planes = c2f(_sb.diag(tt, pres, c2f(np.ascontiguousarray(z)), c2f(np.ascontiguousarray(std)),
                              c2f(np.ascontiguousarray(pick(t))), c2f(np.ascontiguousarray(pick(v))),
                              c2f(np.ascontiguousarray(pick(u))), c2f(np.ascontiguousarray(dist)),
                              c2f(np.ascontiguousarray(ws)), c2f(np.ascontiguousarray(wd)),
                              c2f(np.ascontiguousarray(thc)), **kwargs))
        result[:, rec] = planes
        thc, ws, wd = planes[1], planes[2], planes[3]
        tt += 1
    return tt, result[0], thc, ws, wd


def diag_batched(tt, lsm, z, std, lon, lat, pres, u, v, t, ci=None, ws=None, wd=None, thc=None,
                 target_plev=700.0, thresh_wind=11.0, thresh_winddir=90.0, thresh_windch=5.0, thresh_thc=0.75,
                 target_time=6.0, maxdist=180.0, timestep=24.0, return_flags=False):
    """Same result as ``diag`` for inputs with a time axis, computed by ONE call of the C-ABI
    (sbd_py_diag_loop_f32): the per-record get_edges -> get_dist -> diag sequence of I:219-245 runs
    on the device with the state threaded between records there, and the coast distance is only
    recomputed for records whose ice field changed.  u, v: (time, pres, lat, lon); t, ci: (time, lat,
    lon); arrays may be numpy (host) or torch CUDA tensors (float32, C-contiguous).

    Returns tt, sb_con (time, lat, lon), thc(=t0 of the last record), ws, wd [, flags]."""
    tt = max(1, int(tt))
    f32 = lambda a: a if hasattr(a, "data_ptr") else np.ascontiguousarray(a, dtype=np.float32)
    lsm, z, std, u, v, t = (f32(a) for a in (lsm, z, std, u, v, t))
    lon, lat, pres = (np.ascontiguousarray(a, dtype=np.float32) for a in (lon, lat, pres))
    nrec, nps, nlat, nlon = u.shape
    if tuple(v.shape) != (nrec, nps, nlat, nlon) or tuple(t.shape) != (nrec, nlat, nlon):
        raise ValueError("diag_batched: u, v must be (time, pres, lat, lon) and t (time, lat, lon)")
    if ci is not None:
        ci = f32(_unmask(ci))
        if tuple(ci.shape) != (nrec, nlat, nlon):
            raise ValueError("diag_batched: ci must be (time, lat, lon)")
    ws = _state(ws, lsm, "Windspeed", tt)
    wd = _state(wd, lsm, "Wind direction", tt)
    _state(thc, lsm, "Heating contrast", tt)   # warns like diag(); the kernels never read thc
    ws, wd = f32(ws), f32(wd)
    out = np.empty((nrec, 4, nlat, nlon), dtype=np.float32)
    flags = np.zeros((nrec, nlat, nlon), dtype=np.uint8) if return_flags else None
    lib = _capi.load()
    q = _capi.ptr
    _capi.check(lib.sbd_py_diag_loop_f32(tt, nrec, q(lsm), q(ci), q(lon), q(lat), q(z), q(std), q(pres), q(u), q(v),
                                         q(t), q(ws), q(wd), float(target_plev), float(thresh_wind),
                                         float(thresh_winddir), float(thresh_windch), float(thresh_thc),
                                         float(target_time), float(maxdist), float(timestep), nps, nlon, nlat,
                                         q(out), q(flags)), "diag_batched")
    res = (tt + nrec, out[:, 0].astype(np.float64), out[-1, 1], out[-1, 2], out[-1, 3])
    return res + (flags,) if return_flags else res
