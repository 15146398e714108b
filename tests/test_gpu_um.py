"""UM hook boundary (UM/vn10.7/sea_breeze_diag.F90 = U): the entry points behind the UM's own dummy lists -- two halo
widths, `error`, theta overwritten by t0, get_dist with 2-D latitude / longitude and halo_i x halo_j radii, the
8-argument swap_bounds -- against the oracle, bit for bit.  Runs on the B200 box: -m gpu."""
import numpy as np
import pytest

from seabreeze_param_b200 import synth, um
from util import assert_bits_equal, generic_case, generic_step_inputs

pytestmark = pytest.mark.gpu


def _pad(a, hj, hi):
    out = np.zeros((a.shape[0] + 2 * hj, a.shape[1] + 2 * hi), np.float32)
    out[hj:hj + a.shape[0], hi:hi + a.shape[1]] = a
    return out


def test_um_swap_bounds_two_widths(oracle):
    rng = np.random.default_rng(1)
    nx, ny, hi, hj, lev = 37, 19, 5, 3, 2
    f = rng.random((lev, ny + 2 * hj, nx + 2 * hi)).astype(np.float32)
    ref = f.copy()
    for k in range(lev):
        oracle.um_swap_bounds(ref[k], nx, ny, hi, hj)
    um.swap_bounds(f, nx, ny, lev, hi, hj)
    assert_bits_equal(ref, f, "um swap_bounds")


def test_um_get_edges_and_get_dist(oracle):
    """get_edges on a tdims_l plane with halo_i != halo_j, then get_dist(landfrac, coast) with 2-D true_latitude /
    true_longitude against the oracle's literal scatter (U:547-582 order, halo writes lost at U:584)."""
    cfg = synth.Config("um", 120, 70, 3, 4, "generic", 5.0, False, 1440.0, "um test")
    land, ice, lon, lat, z, sig = synth.static_fields(cfg, seed=4)
    hi, hj = 7, 5
    # a slightly rotated grid: latitude and longitude both vary along both axes
    lon2d = (lon[None, :] * 0.6 + 0.02 * np.arange(cfg.ny)[:, None] + 150.0).astype(np.float32)   # crosses 180
    lat2d = (lat[:, None] * 0.4 + 0.01 * np.arange(cfg.nx)[None, :]).astype(np.float32)
    # get_edges: oracle = generic get_edges on the square part of the halo, compared on the interior + own swap
    mask_g = np.zeros((cfg.ny + 2 * hj, cfg.nx + 2 * hi), np.float32)
    um.get_edges(mask_g, ice, land, hi, hj)
    ref_sq, _ = oracle.get_edges(ice, land, 3)
    ref = _pad(ref_sq[3:-3, 3:-3], hj, hi)
    oracle.um_swap_bounds(ref, cfg.nx, cfg.ny, hi, hj)
    assert_bits_equal(ref, mask_g, "um get_edges")
    land_l = _pad(land, hj, hi)
    ref_d = oracle.um_get_dist(land, mask_g, lat2d, lon2d, hi, hj)
    um.get_dist(land_l, mask_g, lat2d, lon2d, hi, hj)
    assert_bits_equal(ref_d, mask_g, "um get_dist")
    inner = mask_g[hj:-hj, hi:-hi]
    assert (inner == 12000.0).any() and (np.abs(inner) < 180.0).any() and (inner < 0).any()


@pytest.mark.parametrize("halo_s,halo_l", [(4, 7), (6, 4)])
def test_um_seabreeze_diag_two_halo_widths(oracle, halo_s, halo_l):
    """theta / z / sigma with halo_s, mask with halo_l: the result is the oracle's generic trigger with the UM's
    early-exit level search on planes cropped to min(halo_s, halo_l); theta comes back as t0 (U:210-211); error = 0."""
    h = min(halo_s, halo_l)
    cfg = synth.Config("um2", 96, 64, 12, h, "generic", 4.0, False, 1440.0, "um halo test")
    case = generic_case(cfg, oracle)
    ny, nx = cfg.ny, cfg.nx

    def widen(a_h, hh):      # halo h plane -> halo hh plane with junk beyond h (must never be read by the window search)
        out = np.full((ny + 2 * hh, nx + 2 * hh), 7.0e30, np.float32)
        out[hh - h:hh - h + ny + 2 * h, hh - h:hh - h + nx + 2 * h] = a_h
        return out

    so = [np.zeros((ny, nx), np.float32) for _ in range(4)]
    sg = [np.zeros((ny, nx), np.float32) for _ in range(4)]
    _, st = oracle.sigmoid(case["sigma"])
    for step in (1, 2, 15):
        theta_h, p, u, v = generic_step_inputs(cfg, case, step, oracle)
        fo = oracle.seabreeze_diag(cfg.timestep, step, p, u, v, theta_h, case["mask_h"], case["z_h"], case["sigma_h"],
                                   *so, h, level_scan=1)
        th_s = widen(theta_h, halo_s)
        z_s, sig_s = widen(case["z_h"], halo_s), widen(case["sigma_h"], halo_s)
        z_s[z_s > 1e30] = 0.0; sig_s[sig_s > 1e30] = 0.0       # t0 is evaluated over the whole array: keep it finite
        fg = np.zeros((ny, nx), np.uint8)
        err = um.seabreeze_diag(cfg.timestep, step, p, u, v, th_s, z_s, sig_s, widen(case["mask_h"], halo_l), *sg,
                                halo_s=halo_s, halo_l=halo_l, flags=fg)
        assert err == 0
        for a, b, nm in zip(so, sg, ("windspeed", "winddir", "thc", "sb_con")):
            assert_bits_equal(a, b, "%s step %d" % (nm, step))
        assert np.array_equal(fo, fg)
        # theta := t0 in place (U:210-211), statistics over the interior
        d = slice(halo_s - h, halo_s - h + ny + 2 * h)
        dx = slice(halo_s - h, halo_s - h + nx + 2 * h)
        # t0 on the halo-h part: theta_h - (gmma z) sigmoid(sigma) with the interior's statistics.  Checked to a
        # tolerance here (float64 exp): the bit-exact t0 is already inside the trigger results above
        nstd, r = -st[1], st[2]
        t0 = theta_h - (np.float32(-0.0060956) * case["z_h"]) / (1.0 + np.exp((nstd * (case["sigma_h"] - r)).astype(np.float64)))
        np.testing.assert_allclose(th_s[d, dx], t0.astype(np.float32), rtol=1e-6, atol=1e-4)
    # error = 1 for a model without levels (U:198-201)
    assert um.seabreeze_diag(cfg.timestep, 1, np.zeros((0, ny, nx), np.float32), np.zeros((0, ny, nx), np.float32),
                             np.zeros((0, ny, nx), np.float32), th_s, z_s, sig_s, widen(case["mask_h"], halo_l), *sg,
                             halo_s=halo_s, halo_l=halo_l) == 1
