"""ctypes binding of include/sbd.h.  Fails loudly when the CUDA library is missing: there is no
CPU fallback anywhere in this package."""
import ctypes as C
import os

import numpy as np

from . import build as _build

_LIB = None

SBD_VARIANT_GENERIC = 0
SBD_VARIANT_PYTHON = 1

# every symbol declared in include/sbd.h (tests check that the library exports all of them)
SYMBOLS = [
    "sbd_default_params", "sbd_last_error", "sbd_device_count",
    "sbd_get_edges_f32", "sbd_get_dist_f32", "sbd_sigmoid_f32", "sbd_seabreeze_diag_f32",
    "sbd_swap_bounds_f32", "sbd_swap_bounds_ew_f32", "sbd_swap_bounds_dev_f32",
    "sbd_py_get_edges_f32", "sbd_py_get_dist_f32", "sbd_py_sigmoid_f32", "sbd_py_diag_f32", "sbd_py_diag_loop_f32",
    "sbd_get_threads",
    "sbd_plan_create", "sbd_plan_destroy", "sbd_plan_set_stream", "sbd_plan_use_own_stream", "sbd_plan_get_stream",
    "sbd_plan_sync", "sbd_plan_set_static", "sbd_plan_info", "sbd_plan_n_quads", "sbd_plan_host_split", "sbd_plan_step", "sbd_plan_step_scan", "sbd_plan_step_finish", "sbd_plan_step_py",
    "sbd_launch_count", "sbd_stats_chain_f32", "sbd_plan_set_timing", "sbd_plan_last_timing", "sbd_plan_set_halo_mode", "sbd_plan_set_peer_theta",
    "sbd_plan_set_peer_sync", "sbd_plan_set_strip_cells", "sbd_plan_set_tile_split", "sbd_plan_set_tile_rows", "sbd_plan_set_xplane", "sbd_set_stats_mode", "sbd_plan_set_window_kernel", "sbd_set_window_kernel_default", "sbd_plan_set_row_window", "sbd_get_edges_ex_f32", "sbd_plan_window_info",
    "sbd_get_edges_f64", "sbd_get_dist_f64", "sbd_sigmoid_f64", "sbd_seabreeze_diag_f64", "sbd_swap_bounds_f64",
    "sbd_plan_cache_clear", "sbd_plan_cache_stats", "sbd_host_register", "sbd_host_unregister",
    "sbd_um_seabreeze_diag_f32", "sbd_um_get_edges_f32", "sbd_um_get_dist_f32", "sbd_um_swap_bounds_f32",
    "sbd_um_seabreeze_diag_f64", "sbd_um_get_edges_f64", "sbd_um_get_dist_f64", "sbd_um_swap_bounds_f64",
]


class SbdParams(C.Structure):
    _fields_ = [
        ("target_plev", C.c_float), ("thresh_wind", C.c_float), ("thresh_winddir", C.c_float),
        ("thresh_windch", C.c_float), ("thresh_thc", C.c_float), ("target_time", C.c_float),
        ("maxdist", C.c_float), ("gmma", C.c_float), ("rad2deg", C.c_float), ("missing", C.c_float),
        ("level_scan", C.c_int),
    ]


class SbdError(RuntimeError):
    pass


def lib_path():
    return _build.OUT


def load():
    """Load (building first if the sources are newer) the C-ABI shared library."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.OUT
    if _build.needs_build():
        try:
            _build.build()
        except Exception as e:  # no nvcc on this machine and no prebuilt .so
            if not os.path.exists(path):
                raise SbdError(
                    "libsbd_b200.so is missing and cannot be built (%s). The sea-breeze trigger has "
                    "no CPU fallback: build it with `python -m seabreeze_param_b200.build`." % e)
            # a prebuilt library older than its sources: usable (a box without nvcc), but say so -- results of a stale
            # build are not the results of the code next to it
            import warnings
            warnings.warn("libsbd_b200.so is older than its sources and could not be rebuilt (%s); loading the stale "
                          "library" % str(e).splitlines()[0], RuntimeWarning)
    lib = C.CDLL(path)
    vp, fp, ip, i, f = C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_float
    lib.sbd_default_params.argtypes = [C.POINTER(SbdParams), i]
    lib.sbd_default_params.restype = None
    lib.sbd_last_error.restype = C.c_char_p
    lib.sbd_device_count.restype = i
    lib.sbd_get_threads.restype = i
    lib.sbd_launch_count.restype = C.c_ulonglong
    lib.sbd_get_edges_f32.argtypes = [fp, fp, fp, i, i, i, i, vp]
    lib.sbd_get_dist_f32.argtypes = [fp, fp, fp, fp, f, fp, i, i, i]
    lib.sbd_sigmoid_f32.argtypes = [fp, fp, i, i, fp]
    lib.sbd_seabreeze_diag_f32.argtypes = [f, i, fp, fp, fp, fp, fp, fp, fp, fp, fp, fp, fp, i, i, i, i,
                                           C.POINTER(SbdParams), vp]
    lib.sbd_swap_bounds_f32.argtypes = [fp, i, i, i]
    lib.sbd_swap_bounds_ew_f32.argtypes = [fp, i, i, i]
    lib.sbd_swap_bounds_dev_f32.argtypes = [fp, i, i, i, i, vp]
    lib.sbd_py_get_edges_f32.argtypes = [fp, fp, i, i, fp]
    lib.sbd_py_get_dist_f32.argtypes = [fp, fp, fp, fp, i, i, f, fp, ip]
    lib.sbd_py_sigmoid_f32.argtypes = [fp, i, i, fp]
    lib.sbd_py_diag_f32.argtypes = [i, fp, fp, fp, fp, fp, fp, fp, fp, fp, fp, f, f, f, f, f, f, f, f, i, i, i,
                                    fp, vp]
    lib.sbd_py_diag_loop_f32.argtypes = [i, i] + [fp] * 12 + [f] * 8 + [i, i, i, fp, vp]
    lib.sbd_plan_create.argtypes = [C.POINTER(vp), i, i, i, i, i, C.POINTER(SbdParams)]
    lib.sbd_plan_destroy.argtypes = [vp]
    lib.sbd_plan_destroy.restype = None
    lib.sbd_plan_set_stream.argtypes = [vp, vp]
    lib.sbd_plan_use_own_stream.argtypes = [vp]
    lib.sbd_plan_get_stream.argtypes = [vp]
    lib.sbd_plan_get_stream.restype = vp
    lib.sbd_plan_sync.argtypes = [vp]
    lib.sbd_plan_set_static.argtypes = [vp, fp, fp, fp, fp]
    lib.sbd_plan_info.argtypes = [vp, ip, ip, fp]
    lib.sbd_plan_n_quads.argtypes = [vp]
    lib.sbd_plan_host_split.argtypes = [vp, ip, ip]
    lib.sbd_plan_step.argtypes = [vp, f, i, fp, fp, fp, fp, fp, fp, fp, fp, vp]
    lib.sbd_plan_step_scan.argtypes = [vp, fp, fp, fp]
    lib.sbd_plan_step_finish.argtypes = [vp, f, i, fp, fp, fp, fp, fp, vp]
    lib.sbd_plan_step_py.argtypes = [vp, i, f, fp, fp, fp, fp, fp, fp, fp, vp]
    lib.sbd_plan_set_halo_mode.argtypes = [vp, i]
    lib.sbd_plan_set_peer_theta.argtypes = [vp, fp, i, fp]
    lib.sbd_plan_set_peer_sync.argtypes = [vp, vp, vp, vp]
    lib.sbd_plan_set_strip_cells.argtypes = [vp, i]
    lib.sbd_plan_set_tile_split.argtypes = [vp, i]
    lib.sbd_plan_set_tile_rows.argtypes = [vp, i]
    lib.sbd_plan_set_xplane.argtypes = [vp, i]
    lib.sbd_set_stats_mode.argtypes = [i]
    lib.sbd_plan_set_window_kernel.argtypes = [vp, i]
    lib.sbd_set_window_kernel_default.argtypes = [i]
    lib.sbd_plan_set_row_window.argtypes = [vp, i, i]
    lib.sbd_get_edges_ex_f32.argtypes = [fp, fp, fp, i, i, i, i, vp, fp, fp]
    lib.sbd_plan_window_info.argtypes = [vp, vp, vp, vp]
    d = C.c_double
    lib.sbd_get_edges_f64.argtypes = [vp, vp, vp, i, i, i, i]
    lib.sbd_get_dist_f64.argtypes = [vp, vp, vp, vp, d, vp, i, i, i]
    lib.sbd_sigmoid_f64.argtypes = [vp, vp, i, i]
    lib.sbd_seabreeze_diag_f64.argtypes = [d, i] + [vp] * 11 + [i, i, i, i, vp, vp]
    lib.sbd_swap_bounds_f64.argtypes = [vp, i, i, i]
    lib.sbd_plan_set_timing.argtypes = [vp, i]
    lib.sbd_plan_last_timing.argtypes = [vp, fp]
    lib.sbd_stats_chain_f32.argtypes = [fp, i, i, i, i, i, fp]
    lib.sbd_plan_cache_stats.argtypes = [ip, ip, ip]
    lib.sbd_host_register.argtypes = [vp, C.c_size_t]
    lib.sbd_host_unregister.argtypes = [vp]
    lib.sbd_um_seabreeze_diag_f32.argtypes = [f, i] + [fp] * 11 + [i, i, i, i, i, i, C.POINTER(SbdParams), vp]
    lib.sbd_um_get_edges_f32.argtypes = [fp, fp, fp, i, i, i, i]
    lib.sbd_um_get_dist_f32.argtypes = [fp, fp, fp, fp, i, i, i, i, f]
    lib.sbd_um_swap_bounds_f32.argtypes = [fp, i, i, i, i, i]
    lib.sbd_um_seabreeze_diag_f64.argtypes = [C.c_double, i] + [vp] * 11 + [i, i, i, i, i, i, C.POINTER(SbdParams), vp]
    lib.sbd_um_get_edges_f64.argtypes = [vp, vp, vp, i, i, i, i]
    lib.sbd_um_get_dist_f64.argtypes = [vp, vp, vp, vp, i, i, i, i, C.c_double]
    lib.sbd_um_swap_bounds_f64.argtypes = [vp, i, i, i, i, i]
    _LIB = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().sbd_last_error().decode(errors="replace")
        raise SbdError("%s failed with status %d: %s" % (what or "sbd call", rc, msg))


def ptr(a):
    """Raw address of a numpy array, torch tensor (host or CUDA), int address, or None."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError("cannot take the address of %r" % type(a))


_REGISTERED = {}   # address -> (bytes, finalizer)


def auto_register(a, min_bytes=1 << 20):
    """Page-lock + device-map a large host numpy array the first time the library sees it, and undo that when the
    array (or the buffer it views) is garbage collected.  No-op for small / non-numpy / non-owning-unknown arrays,
    when there is no device, or with SBD_AUTO_REGISTER=0."""
    import weakref
    if not isinstance(a, np.ndarray) or a.nbytes < min_bytes or os.environ.get("SBD_AUTO_REGISTER", "1") == "0":
        return
    addr = a.ctypes.data
    ent = _REGISTERED.get(addr)
    if ent is not None and ent[0] >= a.nbytes:
        return
    owner = a
    while isinstance(owner.base, np.ndarray):
        owner = owner.base
    if owner.base is not None:      # memory owned by a foreign object (torch pinned tensor, mmap, ...): leave it alone
        return
    lib = load()
    if lib.sbd_device_count() < 1 or lib.sbd_host_register(addr, a.nbytes) != 0:
        return

    def _drop(addr=addr):
        _REGISTERED.pop(addr, None)
        try:
            lib.sbd_host_unregister(addr)
        except Exception:
            pass
    _REGISTERED[addr] = (a.nbytes, weakref.finalize(owner, _drop))


def default_params(variant):
    p = SbdParams()
    load().sbd_default_params(C.byref(p), variant)
    return p
