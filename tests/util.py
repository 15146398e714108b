"""Shared helpers for the parity tests."""
import numpy as np

from seabreeze_param_b200 import synth


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bits_equal(a, b, what=""):
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, a.shape, b.shape)
    ne = bits(a) != bits(b)
    # NaNs with different payloads still count as equal results
    ne &= ~(np.isnan(a) & np.isnan(b))
    if ne.any():
        idx = np.argwhere(ne)[:5]
        raise AssertionError("%s: %d / %d values differ, first at %s: %r vs %r" % (
            what, int(ne.sum()), ne.size, idx.tolist(), a[tuple(idx[0])], b[tuple(idx[0])]))


def generic_case(cfg, oracle, seed=0, fractional=False):
    """Static fields of a generic-variant case run through the oracle's get_edges/get_dist:
    returns dict with halo'd cdist ("mask"), theta-independent planes etc."""
    land, ice, lon, lat, z, sig = synth.static_fields(cfg, seed=seed, fractional=fractional)
    h = cfg.h
    coast_h, m1 = oracle.get_edges(ice, land, h, swap=True)
    cdist = oracle.get_dist(coast_h, land, lon, lat, 180.0, h, literal=False)
    mask_h = synth.pad_halo(cdist, h)
    oracle.swap_bounds(mask_h, h)
    z_h = synth.pad_halo(z, h); oracle.swap_bounds(z_h, h)
    sig_h = synth.pad_halo(sig, h); oracle.swap_bounds(sig_h, h)
    return dict(land=land, ice=ice, lon=lon, lat=lat, z=z, sigma=sig, coast_h=coast_h, m1=m1, cdist=cdist,
                mask_h=mask_h, z_h=z_h, sigma_h=sig_h)


def generic_step_inputs(cfg, case, step, oracle, seed=0):
    theta, p, u, v = synth.step_fields(cfg, step, case["land"], case["z"], case["lat"], seed=seed)
    theta_h = synth.pad_halo(theta, cfg.h)
    oracle.swap_bounds(theta_h, cfg.h)
    return theta_h, p, u, v


import contextlib


@contextlib.contextmanager
def window_kernel(mode):
    """Process-wide form of the window-sum kernel for the duration of a test: 0 automatic, 1 tile
    kernel (k_trigger), 2 strip kernel (k_win4).  Both must give the oracle's bits."""
    from seabreeze_param_b200 import _capi
    lib = _capi.load()
    _capi.check(lib.sbd_set_window_kernel_default(int(mode)), "set_window_kernel_default")
    try:
        yield
    finally:
        lib.sbd_set_window_kernel_default(0)
