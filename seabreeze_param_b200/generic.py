"""Host-side mirror of the Fortran module ``sea_breeze_diag_mod`` (generic/sea_breeze_diag.f90, "G")
for Python callers, plus the device-resident Plan API.

Routine names, argument order and meaning follow G; arrays are passed as numpy arrays (host) or
torch CUDA tensors (device) whose memory is column-major / longitude-fastest, i.e. a C-ordered
array of shape (ny, nx) or (nz, ny, nx) *is* the Fortran array (nx, ny[, nz]).  Halo'd fields have
shape (ny+2h, nx+2h).  Everything runs in csrc/sbd_cuda.cu via include/sbd.h.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import SBD_VARIANT_GENERIC, SBD_VARIANT_PYTHON, SbdParams  # noqa: F401


def _is_torch(a):
    return hasattr(a, "data_ptr")


def _chk(a, shape, name):
    if tuple(a.shape) != tuple(shape):
        raise ValueError("%s has shape %s, expected %s" % (name, tuple(a.shape), tuple(shape)))
    if _is_torch(a):
        import torch
        if a.dtype != torch.float32 or not a.is_contiguous():
            raise ValueError("%s must be a contiguous float32 tensor" % name)
    else:
        if a.dtype != np.float32 or not a.flags.c_contiguous:
            raise ValueError("%s must be a C-contiguous float32 array" % name)
    return a


def get_edges(mask, icefrac, landfrac, halo_size, swap=True, m1=None, px=None, py=None):
    """call get_edges(mask, icefrac, landfrac, halo_size)  (G:273-373).  mask (ny+2h, nx+2h) is
    written in place: interior = coast flag, halos filled by swap_bounds when ``swap``.
    Optional by-products: ``m1`` (uint8 (ny, nx)) the binarised mask, ``px``, ``py`` (float32 (ny, nx))
    the Sobel components of G:350-357."""
    ny, nx = landfrac.shape
    h = int(halo_size)
    _chk(mask, (ny + 2 * h, nx + 2 * h), "mask")
    _chk(icefrac, (ny, nx), "icefrac")
    _chk(landfrac, (ny, nx), "landfrac")
    lib = _capi.load()
    _capi.check(lib.sbd_get_edges_ex_f32(_capi.ptr(mask), _capi.ptr(icefrac), _capi.ptr(landfrac), nx, ny, h,
                                         1 if swap else 0, _capi.ptr(m1), _capi.ptr(px), _capi.ptr(py)), "get_edges")
    return mask


def get_dist(coast, landfrac, lon, lat, maxdist, cdist, halo_size):
    """call get_dist(coast, landfrac, lon, lat, maxdist, cdist, halo_size)  (G:375-455)."""
    ny, nx = landfrac.shape
    h = int(halo_size)
    _chk(coast, (ny + 2 * h, nx + 2 * h), "coast")
    _chk(landfrac, (ny, nx), "landfrac")
    _chk(lon, (nx,), "lon")
    _chk(lat, (ny,), "lat")
    _chk(cdist, (ny, nx), "cdist")
    lib = _capi.load()
    _capi.check(lib.sbd_get_dist_f32(_capi.ptr(coast), _capi.ptr(landfrac), _capi.ptr(lon), _capi.ptr(lat),
                                     float(maxdist), _capi.ptr(cdist), nx, ny, h), "get_dist")
    return cdist


def sigmoid(ary, sm=None, stats=None):
    """call sigmoid(ary, sm)  (G:457-481).  Returns sm; ``stats`` (float32[4]) receives mean, 2/sigma, r, var."""
    ny, nx = ary.shape
    if sm is None:
        sm = np.empty_like(ary) if not _is_torch(ary) else ary.new_empty(ary.shape)
    lib = _capi.load()
    _capi.check(lib.sbd_sigmoid_f32(_capi.ptr(ary), _capi.ptr(sm), nx, ny, _capi.ptr(stats)), "sigmoid")
    return sm


def swap_bounds(field, halo_size, ew_only=False):
    """call swap_bounds(field, halo_size)  (generic/halo_exchange_mod.f90:12-17), single domain:
    east-west periodic, north-south edge-replicated.  Multi-GPU: see seabreeze_param_b200.halo."""
    h = int(halo_size)
    ny, nx = field.shape[0] - 2 * h, field.shape[1] - 2 * h
    lib = _capi.load()
    fn = lib.sbd_swap_bounds_ew_f32 if ew_only else lib.sbd_swap_bounds_f32
    _capi.check(fn(_capi.ptr(field), nx, ny, h), "swap_bounds")
    return field


def seabreeze_diag(timestep, timestep_number, p, u, v, theta, mask, z, sigma, windspeed, winddir, thc,
                   sb_con, params=None, flags=None):
    """call seabreeze_diag(timestep, timestep_number, p, u, v, theta, mask, z, sigma, windspeed,
    winddir, thc, sb_con)  (G:55-271, dummy-argument names of G:55-56; pass by keyword to be
    order-agnostic).  The library keeps the plan of the last few (dimensions, tunables) and redoes the
    static work only when mask, z or sigma changed (content checksum)."""
    nz, ny, nx = p.shape
    h = (theta.shape[1] - nx) // 2
    for name, a in (("u", u), ("v", v)):
        _chk(a, (nz, ny, nx), name)
    for name, a in (("theta", theta), ("mask", mask), ("z", z), ("sigma", sigma)):
        _chk(a, (ny + 2 * h, nx + 2 * h), name)
    for name, a in (("windspeed", windspeed), ("winddir", winddir), ("thc", thc), ("sb_con", sb_con)):
        _chk(a, (ny, nx), name)
    lib = _capi.load()
    q = _capi.ptr
    prm = C.byref(params) if params is not None else None
    for a in (p, u, v, theta, mask, z, sigma, windspeed, winddir, thc, sb_con):
        _capi.auto_register(a)   # large host arrays: page-locked once, then PCIe rate / zero-copy gather
    _capi.check(lib.sbd_seabreeze_diag_f32(float(timestep), int(timestep_number), q(p), q(u), q(v), q(theta),
                                           q(mask), q(z), q(sigma), q(windspeed), q(winddir), q(thc),
                                           q(sb_con), nx, ny, nz, h, prm, q(flags)), "seabreeze_diag")
    return sb_con


class Plan:
    """Device-resident plan: static fields (mask = signed coast distance, z, sigma) are processed
    once by ``set_static``; ``step`` / ``step_py`` run one timestep.  Arguments may be numpy arrays
    (staged, blocking) or torch CUDA tensors (asynchronous on the plan's stream)."""

    def __init__(self, variant, nx, ny, nz, h=0, params=None):
        self.lib = _capi.load()
        self.variant, self.nx, self.ny, self.nz, self.h = variant, nx, ny, nz, (h if variant == SBD_VARIANT_GENERIC else 0)
        self._h = C.c_void_p()
        prm = C.byref(params) if params is not None else None
        _capi.check(self.lib.sbd_plan_create(C.byref(self._h), variant, nx, ny, nz, h, prm), "plan_create")

    def close(self):
        if self._h:
            self.lib.sbd_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream_ptr):
        _capi.check(self.lib.sbd_plan_set_stream(self._h, stream_ptr), "plan_set_stream")

    @property
    def stream(self):
        return self.lib.sbd_plan_get_stream(self._h)

    def set_halo_mode(self, mode):
        """Fold swap_bounds into the kernels' reads: bit0 E-W periodic, bit1 north edge, bit2 south edge."""
        _capi.check(self.lib.sbd_plan_set_halo_mode(self._h, int(mode)), "plan_set_halo_mode")

    def set_peer_theta(self, theta_up, ny_up, theta_dn):
        """Neighbour ranks' halo'd theta planes (peer-mapped tensors / addresses, or None)."""
        _capi.check(self.lib.sbd_plan_set_peer_theta(self._h, _capi.ptr(theta_up), int(ny_up), _capi.ptr(theta_dn)),
                    "plan_set_peer_theta")

    def set_peer_sync(self, sync_local, sync_up, sync_dn):
        """In-kernel post/wait flags for the neighbour reads (int32 tensors / addresses in peer-mapped
        memory, or None); replaces a barrier launch per step.  See sbd_plan_set_peer_sync."""
        q = _capi.ptr
        _capi.check(self.lib.sbd_plan_set_peer_sync(self._h, q(sync_local), q(sync_up), q(sync_dn)),
                    "plan_set_peer_sync")

    def set_strip_cells(self, cells):
        """Cells per strip of the k_win8 window kernel: 4 (default) or 8.  See sbd_plan_set_strip_cells."""
        _capi.check(self.lib.sbd_plan_set_strip_cells(self._h, int(cells)), "plan_set_strip_cells")

    def set_tile_split(self, max_parts):
        """Most CTAs a heavy tile of the k_win8 window kernel is given to (before set_static).  See sbd_plan_set_tile_split."""
        _capi.check(self.lib.sbd_plan_set_tile_split(self._h, int(max_parts)), "plan_set_tile_split")

    def set_xplane(self, on):
        """Strip window kernels stage from the per-step x-plane (default) or from theta / gzsm / code.  See sbd_plan_set_xplane."""
        _capi.check(self.lib.sbd_plan_set_xplane(self._h, 1 if on else 0), "plan_set_xplane")

    def set_tile_rows(self, rows):
        """Rows per tile of the k_win8 window kernel: 0 auto (default), 16 or 32.  See sbd_plan_set_tile_rows."""
        _capi.check(self.lib.sbd_plan_set_tile_rows(self._h, int(rows)), "plan_set_tile_rows")

    def window_info(self):
        """(nn, n_land, n_sea): int32 (ny, nx) arrays -- the radius at which the growing-window loop
        (G:188-214) stops for every cell and its land / sea counters; 0 outside the coastal zone."""
        out = [np.zeros((self.ny, self.nx), dtype=np.int32) for _ in range(3)]
        _capi.check(self.lib.sbd_plan_window_info(self._h, *[o.ctypes.data for o in out]), "plan_window_info")
        return out

    def set_row_window(self, row_lo, row_hi=-1):
        """Process only rows [row_lo, row_hi) (before set_static).  See sbd_plan_set_row_window."""
        _capi.check(self.lib.sbd_plan_set_row_window(self._h, int(row_lo), int(row_hi)), "plan_set_row_window")

    def set_window_kernel(self, mode):
        """0 auto, 1 tile kernel (k_trigger), 2 k_win4, 3 k_win8.  See sbd_plan_set_window_kernel."""
        _capi.check(self.lib.sbd_plan_set_window_kernel(self._h, int(mode)), "plan_set_window_kernel")

    def use_own_stream(self):
        _capi.check(self.lib.sbd_plan_use_own_stream(self._h), "plan_use_own_stream")

    def set_timing(self, on):
        _capi.check(self.lib.sbd_plan_set_timing(self._h, 1 if on else 0), "plan_set_timing")

    def last_timing(self):
        """(window kernel, scan kernel, join+final kernel) ms of the last step; needs set_timing(True)."""
        ms = np.zeros(3, dtype=np.float32)
        _capi.check(self.lib.sbd_plan_last_timing(self._h, _capi.ptr(ms)), "plan_last_timing")
        return float(ms[0]), float(ms[1]), float(ms[2])

    def sync(self):
        _capi.check(self.lib.sbd_plan_sync(self._h), "plan_sync")

    def set_static(self, mask, z, sigma, stats=None):
        q = _capi.ptr
        st = None
        if stats is not None:
            st = np.ascontiguousarray(stats, dtype=np.float32)
        _capi.check(self.lib.sbd_plan_set_static(self._h, q(mask), q(z), q(sigma), q(st)), "plan_set_static")

    def info(self):
        nn, nc = C.c_int(), C.c_int()
        st = np.zeros(4, dtype=np.float32)
        _capi.check(self.lib.sbd_plan_info(self._h, C.byref(nn), C.byref(nc), _capi.ptr(st)), "plan_info")
        ys, qn = C.c_int(), C.c_int()
        _capi.check(self.lib.sbd_plan_host_split(self._h, C.byref(ys), C.byref(qn)), "plan_host_split")
        return {"nn_max": nn.value, "n_coastal": nc.value, "n_quads": int(self.lib.sbd_plan_n_quads(self._h)),
                "host_split_row": ys.value, "n_quads_north": qn.value, "mean": float(st[0]), "std2inv": float(st[1]),
                "r": float(st[2]), "var": float(st[3])}

    def step(self, timestep, timestep_number, p, u, v, theta, windspeed, winddir, thc, sb_con, flags=None):
        q = _capi.ptr
        _capi.check(self.lib.sbd_plan_step(self._h, float(timestep), int(timestep_number), q(p), q(u), q(v),
                                           q(theta), q(windspeed), q(winddir), q(thc), q(sb_con), q(flags)),
                    "plan_step")

    def step_scan(self, p, u, v):
        """First half of step(): the level scan over p (needs neither theta nor halos)."""
        q = _capi.ptr
        _capi.check(self.lib.sbd_plan_step_scan(self._h, q(p), q(u), q(v)), "plan_step_scan")

    def step_finish(self, timestep, timestep_number, theta, windspeed, winddir, thc, sb_con, flags=None):
        """Second half of step(): window sums on theta (halos valid), thresholds, state update."""
        q = _capi.ptr
        _capi.check(self.lib.sbd_plan_step_finish(self._h, float(timestep), int(timestep_number), q(theta),
                                                  q(windspeed), q(winddir), q(thc), q(sb_con), q(flags)),
                    "plan_step_finish")

    def step_py(self, timestep_number, timestep_min, p, u, v, theta, windspeed, winddir, output, flags=None):
        q = _capi.ptr
        _capi.check(self.lib.sbd_plan_step_py(self._h, int(timestep_number), float(timestep_min), q(p), q(u),
                                              q(v), q(theta), q(windspeed), q(winddir), q(output), q(flags)),
                    "plan_step_py")
