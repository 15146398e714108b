"""Generates the committed golden fixtures tests/golden/*.npz.

The reference ships no golden vectors and cannot be run here (Fortran, no compiler), so these
fixtures are produced by the CPU oracle (oracle/sbd_oracle.cpp: literal loops, pinned math) on
small synthetic inputs; inputs AND outputs are stored so that the check does not depend on the
numpy / scipy versions that generated them.  They pin (a) the oracle against accidental change
and (b) the CUDA path on machines where the oracle is not rebuilt.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as O  # noqa: E402
from seabreeze_param_b200 import synth  # noqa: E402

GEN = synth.Config("g_gen", 48, 32, 6, 3, "generic", 3.0, False, 1440.0, "golden: generic 48x32x6, halo 3")
PY = synth.Config("g_py", 360, 181, 4, 0, "python", 4.0, False, 360.0, "golden: python 360x181, 4 levels")


def make_generic():
    cfg = GEN
    O.set_math(1)
    land, ice, lon, lat, z, sig = synth.static_fields(cfg, seed=5, fractional=True)
    h = cfg.h
    coast_h, m1 = O.get_edges(ice, land, h)
    cdist = O.get_dist(coast_h, land, lon, lat, 180.0, h, literal=True)
    mk = O.swap_bounds(synth.pad_halo(cdist, h), h)
    zh = O.swap_bounds(synth.pad_halo(z, h), h)
    sh = O.swap_bounds(synth.pad_halo(sig, h), h)
    st = [np.zeros((cfg.ny, cfg.nx), np.float32) for _ in range(4)]
    out = dict(land=land, ice=ice, lon=lon, lat=lat, z=z, sigma=sig, coast_h=coast_h, m1=m1, cdist=cdist,
               mask_h=mk, z_h=zh, sigma_h=sh, timestep=np.float32(cfg.timestep), h=np.int32(h))
    for step in (1, 2, 15):
        theta, p, u, v = synth.step_fields(cfg, step, land, z, lat, seed=5)
        th = O.swap_bounds(synth.pad_halo(theta, h), h)
        fl = O.seabreeze_diag(cfg.timestep, step, p, u, v, th, mk, zh, sh, *st, h, literal=True)
        out.update({"theta_h_%d" % step: th, "p_%d" % step: p, "u_%d" % step: u, "v_%d" % step: v,
                    "ws_%d" % step: st[0].copy(), "wd_%d" % step: st[1].copy(), "thc_%d" % step: st[2].copy(),
                    "sb_%d" % step: st[3].copy(), "flags_%d" % step: fl})
    np.savez_compressed(os.path.join(HERE, "generic_48x32x6_h3.npz"), **out)


def make_python():
    """Regional 0.5-degree grid around 70N (80 x 31 cells, 3 levels): get_dist radius k = 3."""
    O.set_math(1)
    nx, ny, nps = 80, 31, 3
    lon = (100.0 + 0.5 * np.arange(nx)).astype(np.float32)
    lat = (75.0 - 0.5 * np.arange(ny)).astype(np.float32)
    rng = np.random.Generator(np.random.PCG64(9))
    f = synth._smooth(rng, ny, nx, 3.0, True)
    land = (f > np.quantile(f, 0.6)).astype(np.float32)
    ice = (np.clip((lat.astype(np.float64) - 68.0) / 6.0, 0, 1)[:, None] * (1 - land) * 0.6).astype(np.float32)
    z = (land * (100 + 900 * rng.random((ny, nx)))).astype(np.float32)
    sig = (land * (5 + 200 * rng.random((ny, nx)))).astype(np.float32)
    p = np.array([85000.0, 70000.0, 50000.0], np.float32)
    coast = O.py_get_edges(land, ice)
    cdist = O.py_get_dist(coast, land, lon, lat, literal=True)
    out = dict(lsm=land, ci=ice, lon=lon, lat=lat, z=z, std=sig, coast=coast, cdist=cdist, p=p,
               k=np.int32(O.py_dist_radius(lon, lat)), timestep=np.float32(360.0))
    ws = np.zeros_like(z)
    wd = np.zeros_like(z)
    for step in (1, 2):
        theta = (285 + 4 * land * np.cos(step) + 0.7 * rng.standard_normal((ny, nx))).astype(np.float32)
        u = (4 * rng.standard_normal((nps, ny, nx))).astype(np.float32)
        v = (4 * rng.standard_normal((nps, ny, nx))).astype(np.float32)
        o, fl = O.py_diag(step, p, z, sig, theta, v, u, cdist, ws, wd, timestep=360.0, literal=True)
        out.update({"theta_%d" % step: theta, "u_%d" % step: u, "v_%d" % step: v,
                    "out_%d" % step: o, "flags_%d" % step: fl})
        ws, wd = o[2].copy(), o[3].copy()
    np.savez_compressed(os.path.join(HERE, "python_80x31x3.npz"), **out)


def make_math():
    """Known-answer table of the pinned transcendental functions (bit patterns)."""
    import ctypes as C
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-20, 20, 500), rng.uniform(-1e-3, 1e-3, 100), [0.0, 1.0, -1.0, 88.0, -100.0]]).astype(np.float32)
    y = rng.uniform(-30, 30, x.size).astype(np.float32)
    src = r'''
    #include "sbd_math.h"
    extern "C" void tab(const float* x, const float* y, int n, float* e, float* s, float* c, float* a2) {
      for (int i = 0; i < n; ++i) { e[i] = sbd_expf(x[i]); s[i] = sbd_sinf(x[i]); c[i] = sbd_cosf(x[i]); a2[i] = sbd_atan2f(y[i], x[i]); } }
    '''
    import subprocess, tempfile
    d = tempfile.mkdtemp()
    open(os.path.join(d, "t.cpp"), "w").write(src)
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(d, "t.cpp"), "-o", os.path.join(d, "t.so")])
    L = C.CDLL(os.path.join(d, "t.so"))
    e, s, c, a2 = (np.empty_like(x) for _ in range(4))
    L.tab(x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), C.c_int(x.size), *(a.ctypes.data_as(C.c_void_p) for a in (e, s, c, a2)))
    np.savez_compressed(os.path.join(HERE, "pinned_math.npz"), x=x, y=y, expf=e, sinf=s, cosf=c, atan2f=a2)


if __name__ == "__main__":
    make_generic()
    make_python()
    make_math()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
