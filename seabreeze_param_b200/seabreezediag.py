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
