"""Synthetic fields for BASELINE.json's configurations.

The reference ships no data (python_wrapper/run.conf points at ERA-Interim files under ~/Data that
do not exist), so every test and benchmark uses these generators.  All arrays are float32,
C-ordered (ny, nx) / (nz, ny, nx) -- byte-identical to Fortran (nx, ny[, nz]) column-major arrays.
Everything is derived from numpy's PCG64 with fixed seeds so the oracle and the GPU read identical
bytes.
"""
import math
from dataclasses import dataclass

import numpy as np

try:  # scipy is in the image; only the mask smoothing needs it
    from scipy.ndimage import gaussian_filter
except Exception:  # pragma: no cover
    gaussian_filter = None


@dataclass(frozen=True)
class Config:
    name: str
    nx: int
    ny: int
    nz: int          # levels of p,u,v (python variant: nps)
    h: int           # halo / get_dist radius for the generic variant
    variant: str     # "generic" | "python"
    filt: float      # Gaussian filter width (cells) of the land-sea mask
    regional: bool
    timestep: float  # seconds (generic) / minutes (python)
    descr: str


# generic/get_all_fields_mod.f90:9-11 (C1), BASELINE.json configs[1..4] (C2..C5)
CONFIGS = {
    "c1": Config("c1", 128, 96, 56, 4, "generic", 3.0, False, 1440.0,
                 "dummy_model default grid 128x96x56, halo 4 (generic/get_all_fields_mod.f90:9-10)"),
    "c2": Config("c2", 1024, 769, 70, 6, "generic", 8.0, False, 1440.0,
                 "global N512-class grid 1024x769x70, halo 6, one timestep"),
    "c3": Config("c3", 4000, 3000, 70, 54, "generic", 40.0, True, 1440.0,
                 "regional 2.5 km tropical coastal domain 4000x3000x70, halo 54"),
    "c4": Config("c4", 3600, 1800, 37, 0, "python", 30.0, False, 60.0,
                 "global 0.1 degree grid 3600x1800, 37 pressure levels, 24 hourly records"),
    "c5": Config("c5", 480, 241, 37, 0, "python", 4.0, False, 360.0,
                 "ERA-Interim-like 0.75 degree grid 480x241, 37 levels (python_wrapper/run.conf:5)"),
    "c3s": Config("c3s", 1000, 752, 4, 54, "generic", 40.0, True, 1440.0,
                  "probe: 1/16 of the regional 2.5 km domain (same mask statistics as c3), 4 levels"),
    "c3r8": Config("c3r8", 4000, 384, 4, 54, "generic", 40.0, True, 1440.0,
                   "probe: one of eight row blocks of the regional 2.5 km domain (a rank's share at 8 GPUs), 4 levels"),
    # small cases for unit tests
    "t_gen": Config("t_gen", 96, 64, 12, 5, "generic", 4.0, False, 1440.0, "test: small generic grid"),
    "t_py": Config("t_py", 360, 181, 9, 0, "python", 4.0, False, 360.0, "test: small python-variant grid"),
}

ERA_LEVELS_HPA = [1, 2, 3, 5, 7, 10, 20, 30, 50, 70, 100, 125, 150, 175, 200, 225, 250, 300, 350, 400,
                  450, 500, 550, 600, 650, 700, 750, 775, 800, 825, 850, 875, 900, 925, 950, 975, 1000]


def _smooth(rng, ny, nx, width, periodic_x):
    w = rng.standard_normal((ny, nx))
    if gaussian_filter is None:
        raise RuntimeError("scipy is required for the synthetic mask generator")
    f = gaussian_filter(w, width, mode=("nearest", "wrap" if periodic_x else "nearest"))
    f -= f.mean()
    s = f.std()
    return f / (s if s > 0 else 1.0)


def grid(cfg):
    """lon(nx), lat(ny) in degrees.  Latitude is descending (north first) so that the python
    get_dist's 70-degree row is never the last row."""
    if cfg.regional:
        lon = 90.0 + 0.0225 * np.arange(cfg.nx)
        lat = 10.0 - 0.0225 * np.arange(cfg.ny)
        lat = lat + 0.0  # 2.5 km spacing around the equator
    else:
        lon = np.arange(cfg.nx) * (360.0 / cfg.nx)
        lat = 90.0 - np.arange(cfg.ny) * (180.0 / (cfg.ny - 1))
    return lon.astype(np.float32), lat.astype(np.float32)


def static_fields(cfg, seed=0, fractional=False):
    """landfrac, icefrac, lon, lat, z, sigma (all float32)."""
    lon, lat = grid(cfg)
    per = not cfg.regional
    f = _smooth(np.random.Generator(np.random.PCG64(seed)), cfg.ny, cfg.nx, cfg.filt, per)
    thr = np.quantile(f, 0.70)
    if fractional:
        land = np.clip(0.5 + (f - thr) * 2.0, 0.0, 1.0)
    else:
        land = (f > thr).astype(np.float64)
    if cfg.regional:
        ice = np.zeros_like(land)
    else:
        ice = np.clip((np.abs(lat.astype(np.float64)) - 65.0) / 10.0, 0.0, 1.0)[:, None] * (1.0 - land)
    n1 = _smooth(np.random.Generator(np.random.PCG64(seed + 1)), cfg.ny, cfg.nx, max(cfg.filt / 2, 1.0), per)
    n2 = np.random.Generator(np.random.PCG64(seed + 2)).random((cfg.ny, cfg.nx))
    z = land * (200.0 + 1800.0 * np.clip(0.5 + 0.25 * n1, 0.0, 1.0))
    sigma = land * (5.0 + 300.0 * n2)
    a = np.float32
    return (land.astype(a), ice.astype(a), lon, lat, z.astype(a), sigma.astype(a))


def pressure_1d(nps):
    """ERA-style pressure levels in Pa (python variant takes a 1-D vector, P:76-77)."""
    lev = np.array(ERA_LEVELS_HPA, dtype=np.float64)
    if nps != len(lev):
        lev = np.geomspace(1.0, 1000.0, nps)
    return (lev * 100.0).astype(np.float32)


def step_fields(cfg, step, land, z, lat, seed=0, hour=None):
    """theta (ny,nx), p, u, v for one timestep.  generic: p,u,v are (nz,ny,nx); python: p is 1-D."""
    ny, nx, nz = cfg.ny, cfg.nx, cfg.nz
    per = not cfg.regional
    if hour is None:
        dt_h = (cfg.timestep / 3600.0) if cfg.variant == "generic" else (cfg.timestep / 60.0)
        hour = step * dt_h
    rng = np.random.Generator(np.random.PCG64(seed + 1000 + step))
    latr = np.deg2rad(lat.astype(np.float64))[:, None]
    theta = (300.0 - 30.0 * np.sin(latr) ** 2 + 3.0 * land * math.cos(2 * math.pi * hour / 24.0)
             + 0.5 * rng.standard_normal((ny, nx)))
    base_u = 6.0 * _smooth(np.random.Generator(np.random.PCG64(seed + 3)), ny, nx, max(cfg.filt, 2.0), per)
    base_v = 6.0 * _smooth(np.random.Generator(np.random.PCG64(seed + 4)), ny, nx, max(cfg.filt, 2.0), per)
    ang = math.radians(10.0 * hour)
    ru = math.cos(ang) * base_u - math.sin(ang) * base_v
    rv = math.sin(ang) * base_u + math.cos(ang) * base_v
    lev_scale = (0.6 + 0.8 * np.arange(nz) / max(nz - 1, 1))[:, None, None]
    u = ru[None] * lev_scale + rng.standard_normal((nz, ny, nx))
    v = rv[None] * lev_scale + rng.standard_normal((nz, ny, nx))
    if cfg.variant == "generic":
        k = np.arange(nz, dtype=np.float64)[:, None, None]
        p = 101325.0 * np.exp(-z.astype(np.float64) / 8000.0)[None] * np.exp(-k * math.log(100.0) / max(nz - 1, 1))
        p = p.astype(np.float32)
    else:
        p = pressure_1d(nz)
    a = np.float32
    return theta.astype(a), p, u.astype(a), v.astype(a)


def pad_halo(a, h):
    """Copy an (ny,nx) field into the interior of a zeroed (ny+2h, nx+2h) array."""
    ny, nx = a.shape
    out = np.zeros((ny + 2 * h, nx + 2 * h), dtype=a.dtype)
    out[h:h + ny, h:h + nx] = a
    return out
