#!/usr/bin/env python
"""Golden vectors produced by THE REFERENCE'S OWN SOURCE, executed here.

The reference cannot be compiled in this image (no Fortran compiler), so its kernels are run by
tests/golden/f90interp.py, an interpreter for the Fortran subset they are written in (binary32 arithmetic, one
IEEE operation per Fortran operation, platform libm through ctypes -- see its header).  The .f90 files are read
from /root/reference UNMODIFIED at generation time; only inputs and outputs are stored (tests/golden/ref_*.npz),
together with the sha256 of each source file that produced them.  tests/test_reference_golden.py replays the
fixtures against the oracle (CPU, any machine) and tests/test_gpu_r2.py against the CUDA path (B200).

Cases (all through the reference's python-variant files, the only complete, self-consistent variant):
  ref_py_a   40 x 24 cells at 0.75 degree: get_edges, get_dist (k = 2), then seabreezediag.diag()'s call sequence
             (python_wrapper/seabreezediag/__init__.py:219-245) for records 1, 2, 3 with time-varying sea ice;
             record 3 is a 6-hourly step (timestep 120 min).
  ref_py_b   48 x 20 cells, 2.25 x 0.75 degree boxes, maxdist = 120 km: the search window is wider than
             2*maxdist, so the visit-order reset of sobel.f90:188 decides values; thresholds changed from
             their defaults; first longitude 0, last 105.75 (column-wrap quirk of sobel.f90:67-68 in play).
  ref_py_c   ragged tiny grids (3 x 2, 5 x 4, 9 x 3) for get_edges' clamped / wrapped neighbour indices.
  ref_sigmoid  sigmoid() of the python file and of generic/sea_breeze_diag.f90 (assumed-shape version) on a
             field with zeros, ties and a wide dynamic range.
Run:  python tests/golden/make_reference_golden.py     (needs /root/reference; takes ~1 minute)
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, HERE]
import f90interp as F  # noqa: E402
from seabreeze_param_b200 import synth  # noqa: E402

REF = "/root/reference"
SRC = {
    "sobel": os.path.join(REF, "python_wrapper/seabreezediag/sobel.f90"),
    "diag": os.path.join(REF, "python_wrapper/seabreezediag/seabreeze_diag_python.f90"),
    "generic": os.path.join(REF, "generic/sea_breeze_diag.f90"),
}
f32 = np.float32


def Fo(a):
    """C-ordered (ny, nx[, ...]) -> Fortran-shaped (nx, ny[, ...]) array owning its memory in Fortran order."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    return np.asfortranarray(a.transpose(tuple(range(a.ndim))[::-1]))


def Co(a):
    return np.ascontiguousarray(a.transpose(tuple(range(a.ndim))[::-1]))


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def run_edges(M, lsm, ci):
    ny, nx = lsm.shape
    coast = np.full((nx, ny), np.nan, f32, order="F")
    M.call("get_edges", Fo(lsm), Fo(ci), nx, ny, coast)
    assert np.isfinite(coast).all()
    return Co(coast)


def run_dist(M, coast, lsm, lon, lat, maxdist):
    ny, nx = lsm.shape
    cd = np.full((nx, ny), np.nan, f32, order="F")
    M.call("get_dist", Fo(coast), Fo(lsm), lon.astype(f32).copy(), lat.astype(f32).copy(), nx, ny, f32(maxdist), cd)
    assert np.isfinite(cd).all()
    return Co(cd)


def run_diag(M, step, p, z, std, theta, v, u, cdist, ws, wd, thc, kw):
    """One call of diag(); returns output (4, ny, nx) with NaN where the reference wrote nothing."""
    ny, nx = z.shape
    out = np.full((nx, ny, 4), np.nan, f32, order="F")
    M.allow_nan = True
    M.call("diag", int(step), p.astype(f32).copy(), Fo(z), Fo(std), Fo(theta), Fo(v), Fo(u), Fo(cdist), Fo(ws), Fo(wd),
           Fo(thc), f32(kw["target_plev"]), f32(kw["thresh_wind"]), f32(kw["thresh_winddir"]), f32(kw["thresh_windch"]),
           f32(kw["thresh_thc"]), f32(kw["target_time"]), f32(kw["maxdist"]), f32(kw["timestep"]), len(p), nx, ny, out)
    return Co(out)


DEFAULTS = dict(target_plev=700.0, thresh_wind=11.0, thresh_winddir=90.0, thresh_windch=5.0, thresh_thc=0.75,
                target_time=6.0, maxdist=180.0, timestep=24.0)


def time_loop_case(M, name, nx, ny, dlon, dlat, nps, nrec, kw, dist_maxdist, seed, lat0=66.0):
    cfg = synth.Config(name, nx, ny, nps, 0, "python", 3.0, False, kw["timestep"], name)
    land, ice, _, _, z, sig = synth.static_fields(cfg, seed=seed)
    lon = (np.arange(nx) * dlon).astype(f32)
    lat = (lat0 + 10.0 - np.arange(ny) * dlat).astype(f32)      # descending, crosses 70 N early (S:129-132)
    recs = dict(theta=[], u=[], v=[], ci=[], coast=[], cdist=[], out=[])
    ws = np.zeros_like(z); wd = np.zeros_like(z); thc = np.zeros_like(z)
    p = None
    for r in range(nrec):
        theta, p, u, v = synth.step_fields(cfg, r + 1, land, z, lat, seed=seed)
        ci = np.clip(ice + 0.15 * r * (1.0 - land) * (lat[:, None] > 68.0), 0, 1).astype(f32)
        coast = run_edges(M, land, ci)                                             # I:228
        cd = run_dist(M, coast, land, lon, lat, dist_maxdist)                      # I:228 (maxdist: get_dist's own)
        out = run_diag(M, 1 + r, p, z, sig, theta, v, u, cd, ws, wd, thc, kw)      # I:229-239
        for k, a in (("theta", theta), ("u", u), ("v", v), ("ci", ci), ("coast", coast), ("cdist", cd), ("out", out)):
            recs[k].append(a)
        # I:244: thc, ws, wd = out[1], out[2], out[3]; the last row of out is never written by the reference
        # (P:165) -- the wrapper feeds back whatever f2py's uninitialised output holds there; here: the inputs
        nxt = [np.where(np.isnan(out[k]), prev, out[k]).astype(f32) for k, prev in ((1, thc), (2, ws), (3, wd))]
        thc, ws, wd = nxt
    return dict(land=land, z=z, sig=sig, lon=lon, lat=lat, p=p.astype(f32), dist_maxdist=f32(dist_maxdist),
                **{k: np.stack(v) for k, v in recs.items()}, **{"kw_" + k: f32(v) for k, v in kw.items()})


def main():
    if not os.path.isdir(REF):
        raise SystemExit("needs the reference tree at /root/reference")
    texts = {k: open(v).read() for k, v in SRC.items()}
    M = F.Machine(texts["sobel"], texts["diag"])
    meta = {"sha256_" + k: np.array(sha(v)) for k, v in SRC.items()}
    meta["how"] = np.array("reference .f90 sources executed by tests/golden/f90interp.py (binary32, platform libm)")

    kw = dict(DEFAULTS, timestep=120.0)
    a = time_loop_case(M, "ref_py_a", 40, 24, 0.75, 0.75, 5, 3, kw, 180.0, seed=11)
    np.savez_compressed(os.path.join(HERE, "ref_py_a.npz"), **a, **meta)
    print("ref_py_a: statements executed so far", M.stmt_count, "trigger cells per record",
          [int(((o[0] != 0) & (o[0] < 1e19)).sum()) for o in a["out"]])

    kw = dict(DEFAULTS, timestep=360.0, maxdist=120.0, thresh_wind=9.0, thresh_thc=0.5, thresh_winddir=60.0,
              thresh_windch=4.0, target_plev=850.0)
    b = time_loop_case(M, "ref_py_b", 48, 20, 2.25, 0.75, 4, 2, kw, 120.0, seed=23, lat0=62.0)
    np.savez_compressed(os.path.join(HERE, "ref_py_b.npz"), **b, **meta)
    far = (b["cdist"][0] == 12000.0).sum()
    print("ref_py_b: cells reset / never reached:", int(far), "of", b["cdist"][0].size)

    rng = np.random.default_rng(5)
    c = {}
    for i, (ny, nx) in enumerate(((2, 3), (4, 5), (3, 9))):
        lsm = (rng.random((ny, nx)) > 0.55).astype(f32)
        ci = (rng.random((ny, nx)) * 0.6).astype(f32)
        c["lsm%d" % i], c["ci%d" % i], c["coast%d" % i] = lsm, ci, run_edges(M, lsm, ci)
    np.savez_compressed(os.path.join(HERE, "ref_py_c.npz"), **c, **meta)

    # sigmoid: python file (explicit shape) and generic file (assumed shape, G:457-481)
    MG = F.Machine(texts["generic"])
    ary = (rng.random((37, 53)) * 300).astype(f32)
    ary[::3, ::2] = 0.0
    ary[5, 7] = 3.0e4
    ary[11, :] = 12.5
    nyy, nxx = ary.shape
    sm_p = np.full((nxx, nyy), np.nan, f32, order="F")
    M.call("sigmoid", Fo(ary), nxx, nyy, sm_p)
    sm_g = np.full((nxx, nyy), np.nan, f32, order="F")
    MG.call("sigmoid", Fo(ary), sm_g)
    assert np.array_equal(sm_p.view(np.uint32), sm_g.view(np.uint32)), "python and generic sigmoid differ"
    np.savez_compressed(os.path.join(HERE, "ref_sigmoid.npz"), ary=ary, sm=Co(sm_p), **meta)
    print("done; statements executed:", M.stmt_count + MG.stmt_count)


if __name__ == "__main__":
    main()
