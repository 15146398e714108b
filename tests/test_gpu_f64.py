"""_f64 entry points of the C-ABI (include/sbd.h) for hosts built with promoted reals.

(1) get_edges / get_dist / swap_bounds (static, one-off): twin(x) == widen(f32 routine(narrow(x))) bit for bit -- on
    fp32-representable inputs that is the oracle-checked fp32 result itself;
(2) sigmoid and seabreeze_diag run in BINARY64 THROUGHOUT (csrc/sbd_f64.cuh): against a restatement of
    generic/sea_breeze_diag.f90 in numpy float64 (below: what a -r8 build of the reference computes) the
    diagnostics agree to 1e-9 relative (the two sides differ only in the last ulps of exp / atan2 and in numpy's
    pairwise summation of the statistics), and the trigger flag differs at most where a quantity sits within that of a
    threshold; the UM entry point (two halo widths, theta := t0) is checked the same way."""
import numpy as np
import pytest

from seabreeze_param_b200 import generic, generic64, synth
from util import assert_bits_equal, generic_case, generic_step_inputs

pytestmark = pytest.mark.gpu


def diag_f64(timestep, step, p, u, v, theta, mask, z, sigma, ws, wd, thc, sb, h):
    """generic/sea_breeze_diag.f90:55-271 in float64 (statement order of the source; nn capped at h)."""
    gmma, rad2deg, target_plev, maxdist = -0.0060956, 57.2957, 70000.0, 180.0
    nz, ny, nx = p.shape
    si = sigma[h:h + ny, h:h + nx]
    mean = si.sum() / (nx * ny)                                      # G:466
    std = 2.0 / np.sqrt(((si - mean) ** 2).sum() / (nx * ny))        # G:471-478
    r = (si.max() - si.min()) / 4.0                                  # G:479
    t0 = theta - (gmma * z) * (1.0 / (1.0 + np.exp(-std * (sigma - r))))   # G:480, G:167
    six = np.fmod(step * timestep, 6.0 * 3600.0) < 0.0001            # G:264
    flags = np.zeros((ny, nx), np.uint8)
    for i in range(ny):
        for j in range(nx):
            m = mask[i + h, j + h]
            if abs(m) > maxdist:                                     # G:174-176
                sb[i, j] = 0.0
                continue
            mul = 1.0 if m >= 0.0 else -1.0
            nn = 1
            while True:                                              # G:188-214
                wm = mask[i + h - nn:i + h + nn + 1, j + h - nn:j + h + nn + 1]
                land = wm >= 0.0
                n_l, n_s = int(land.sum()), int((~land).sum())
                if (n_l > 0 and n_s > 0) or nn >= max(h, 1):
                    break
                nn += 1
            wt = t0[i + h - nn:i + h + nn + 1, j + h - nn:j + h + nn + 1]
            with np.errstate(all="ignore"):
                n_thc = mul * (np.float64(wt[land].sum()) / n_l - np.float64(wt[~land].sum()) / n_s)   # G:216
            lev = int(np.argmin(np.abs(p[:, i, j] - target_plev)))   # G:223
            uu, vv = u[lev, i, j], v[lev, i, j]
            n_ws = np.sqrt(uu * uu + vv * vv)                        # G:225
            n_wd = np.arctan2(-uu, -vv) * rad2deg                    # G:227
            if step < 2:
                thc[i, j], wd[i, j], ws[i, j] = n_thc, n_wd, n_ws    # G:235-239
            mws = (ws[i, j] + n_ws) / 2.0
            dws = abs(ws[i, j] - n_ws)
            x = (wd[i, j] - n_wd) + 180.0
            mod = x - np.floor(x / 360.0) * 360.0                    # modulo(x, 360)
            dwd = abs(mod - 180.0)
            fl = 1
            if dwd < 90.0 and dws < 5.0 and mws < 11.0 and abs(n_thc) > 0.75:   # G:249-252
                sb[i, j] = ((abs(n_thc) - 0.75) / n_thc) * ((11.0 - mws) / max(1.0, mws))
                fl = 3
            else:
                sb[i, j] = 0.0
            ws[i, j] = n_ws
            thc[i, j] = n_thc
            if six:
                wd[i, j] = n_wd
            flags[i, j] = fl
    return flags


def test_f64_twins_equal_f32_path_on_representable_inputs(oracle):
    cfg = synth.CONFIGS["t_gen"]
    case = generic_case(cfg, oracle)
    ny, nx, h = cfg.ny, cfg.nx, cfg.h
    d = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    mask64 = np.zeros((ny + 2 * h, nx + 2 * h), np.float64)
    generic64.get_edges(mask64, d(case["ice"]), d(case["land"]), h)
    assert_bits_equal(case["coast_h"], mask64.astype(np.float32), "coast (f64 twin)")
    assert np.array_equal(mask64, mask64.astype(np.float32).astype(np.float64))
    cd64 = np.zeros((ny, nx), np.float64)
    generic64.get_dist(mask64, d(case["land"]), d(case["lon"]), d(case["lat"]), 180.0, cd64, h)
    assert_bits_equal(case["cdist"], cd64.astype(np.float32), "cdist (f64 twin)")
    fld = d(synth.pad_halo(case["z"], h)); ref = synth.pad_halo(case["z"], h).copy()
    generic64.swap_bounds(fld, h); oracle.swap_bounds(ref, h)
    assert_bits_equal(ref, fld.astype(np.float32), "swap_bounds (f64 twin)")
    # sigmoid: binary64 statistics (sequential) and logistic
    sg = d(case["sigma"])
    sm64 = generic64.sigmoid(sg)
    mean = sg.sum() / sg.size
    std = 2.0 / np.sqrt(((sg - mean) ** 2).sum() / sg.size)
    r = (sg.max() - sg.min()) / 4.0
    np.testing.assert_allclose(sm64, 1.0 / (1.0 + np.exp(-std * (sg - r))), rtol=1e-10, atol=1e-14)


def test_f64_twin_vs_64bit_evaluation(oracle):
    cfg = synth.Config("t_f64", 64, 48, 8, 4, "generic", 3.0, False, 1440.0, "f64 tolerance test")
    case = generic_case(cfg, oracle)
    ny, nx, h = cfg.ny, cfg.nx, cfg.h
    rng = np.random.default_rng(5)
    # 64-bit inputs that are NOT fp32-representable: the fp32 fields times (1 + 1e-9 noise)
    w = lambda a: np.ascontiguousarray(a, np.float64) * (1.0 + 1e-9 * rng.standard_normal(a.shape))
    mask64, z64, sg64 = np.ascontiguousarray(case["mask_h"], np.float64), w(case["z_h"]), w(case["sigma_h"])
    st_r = [np.zeros((ny, nx), np.float64) for _ in range(4)]
    st_g = [np.zeros((ny, nx), np.float64) for _ in range(4)]
    n_coast = int((np.abs(case["cdist"]) <= 180.0).sum())
    mism = 0
    for step in (1, 2, 15):
        theta_h, p, u, v = generic_step_inputs(cfg, case, step, oracle)
        th64, p64, u64, v64 = w(theta_h), w(p), w(u), w(v)
        fr = diag_f64(cfg.timestep, step, p64, u64, v64, th64, mask64, z64, sg64, *st_r, h)
        fg = np.zeros((ny, nx), np.uint8)
        generic64.seabreeze_diag(cfg.timestep, step, p64, u64, v64, th64, mask64, z64, sg64, *st_g, flags=fg)
        same = fr == fg
        mism += int((~same).sum())
        np.testing.assert_allclose(st_g[0], st_r[0], rtol=1e-12, atol=1e-12, err_msg="windspeed step %d" % step)
        np.testing.assert_allclose(st_g[2], st_r[2], rtol=1e-9, atol=1e-9, err_msg="thc step %d" % step)
        dwd = np.abs(((st_g[1] - st_r[1]) + 180.0) % 360.0 - 180.0)
        assert dwd.max() < 1e-10, "winddir step %d: %g degrees" % (step, dwd.max())
        np.testing.assert_allclose(st_g[3][same], st_r[3][same], rtol=1e-8, atol=1e-9, err_msg="sb_con step %d" % step)
        # keep both state chains on the same trajectory where a flag differed (threshold rounding)
        for a, b in zip(st_g, st_r):
            a[~same] = b[~same]
    assert mism <= 1, "%d flag mismatches in %d coastal cells" % (mism, n_coast)


def test_um_f64_entry_point_two_halo_widths(oracle):
    """sbd_um_seabreeze_diag_f64: theta / z / sigma with halo 5, mask with halo 7, early-exit level search (U:265-274),
    theta overwritten by t0 (U:210-211) -- against the float64 restatement on planes cropped to halo 5."""
    import ctypes as C
    from seabreeze_param_b200 import _capi
    cfg = synth.Config("t_um64", 56, 40, 8, 5, "generic", 3.0, False, 1440.0, "um f64 test")
    case = generic_case(cfg, oracle)
    ny, nx, hs, hl = cfg.ny, cfg.nx, 5, 7
    d = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    mask_l = np.zeros((ny + 2 * hl, nx + 2 * hl), np.float64)
    mask_l[hl - hs:hl - hs + ny + 2 * hs, hl - hs:hl - hs + nx + 2 * hs] = case["mask_h"]
    st_r = [np.zeros((ny, nx), np.float64) for _ in range(4)]
    st_g = [np.zeros((ny, nx), np.float64) for _ in range(4)]
    lib = _capi.load()
    q = _capi.ptr
    for step in (1, 2):
        theta_h, p, u, v = generic_step_inputs(cfg, case, step, oracle)
        # monotone pressure columns: the early-exit scan and minloc pick the same level
        fr = diag_f64(cfg.timestep, step, d(p), d(u), d(v), d(theta_h), d(case["mask_h"]), d(case["z_h"]), d(case["sigma_h"]),
                      *st_r, hs)
        th = d(theta_h)
        fg = np.zeros((ny, nx), np.uint8)
        p64, u64, v64, z64, sg64 = d(p), d(u), d(v), d(case["z_h"]), d(case["sigma_h"])   # keep the arrays alive
        rc = lib.sbd_um_seabreeze_diag_f64(C.c_double(cfg.timestep), step, q(p64), q(u64), q(v64), q(th), q(z64),
                                           q(sg64), q(mask_l), *[q(a) for a in st_g], nx, ny, cfg.nz, hs, hl, 1,
                                           None, q(fg))
        assert rc == 0
        assert int((fr != fg).sum()) <= 1
        np.testing.assert_allclose(st_g[2], st_r[2], rtol=1e-9, atol=1e-9)
        sg = d(case["sigma"])
        mean = sg.sum() / sg.size
        std = 2.0 / np.sqrt(((sg - mean) ** 2).sum() / sg.size)
        r = (sg.max() - sg.min()) / 4.0
        t0 = d(theta_h) - (-0.0060956 * d(case["z_h"])) * (1.0 / (1.0 + np.exp(-std * (d(case["sigma_h"]) - r))))
        np.testing.assert_allclose(th, t0, rtol=1e-12, atol=1e-10)     # theta := t0 (U:210-211)
