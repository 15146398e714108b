"""Known-answer vectors for the oracle, derived by hand from the reference source (SURVEY.md
Appendix B; the reference itself ships no tests or fixtures)."""
import numpy as np
import pytest

from seabreeze_param_b200 import synth


def test_edges_uniform_and_single_cell(oracle):
    ny, nx = 12, 16
    z = np.zeros((ny, nx), np.float32)
    assert oracle.py_get_edges(z, z).sum() == 0                      # B.1 uniform -> no coast
    one = np.ones((ny, nx), np.float32)
    assert oracle.py_get_edges(one, z).sum() == 0
    lsm = z.copy(); lsm[5, 7] = 1.0
    coast = oracle.py_get_edges(lsm, z)
    exp = np.zeros_like(z); exp[4:7, 6:9] = 1.0; exp[5, 7] = 0.0     # 8 neighbours, not the centre
    assert np.array_equal(coast, exp)


def test_edges_straight_coastline(oracle):
    ny, nx, x0 = 10, 20, 9
    lsm = np.zeros((ny, nx), np.float32); lsm[:, x0:nx - 4] = 1.0    # land for x0 <= x < nx-4
    coast = oracle.py_get_edges(lsm, np.zeros_like(lsm))
    cols = np.where(coast.any(axis=0))[0].tolist()
    assert cols == [x0 - 1, x0, nx - 5, nx - 4]                      # B.2: the two columns at each edge
    assert (coast[:, x0] == 1).all() and (coast[:, x0 - 1] == 1).all()


def test_edges_python_wrap_quirk(oracle):
    """S:68: xx = max(1, modulo(x+j, nlons)): the last column is never read as a neighbour."""
    ny, nx = 8, 16
    lsm = np.zeros((ny, nx), np.float32); lsm[:, nx - 1] = 1.0       # land only in the last column
    coast = oracle.py_get_edges(lsm, np.zeros_like(lsm))
    assert coast.sum() == 0                                          # nobody ever sees it


def test_edges_generic_binarise(oracle):
    land = np.array([[0.6, 0.4, 0.4, 0.1]], np.float32)
    ice = np.array([[0.0, 0.2, 0.3, 0.5]], np.float32)
    _, m1 = oracle.get_edges(ice, land, 1, swap=True)
    # G:325-337: ice<=0.2 -> land>=0.5 ; else land+ice>=0.5
    assert m1.tolist() == [[1, 0, 1, 1]]


def test_dist_self_far_and_sign(oracle):
    cfg = synth.CONFIGS["c5"]
    lon, lat = synth.grid(cfg)
    ny, nx = cfg.ny, cfg.nx
    lsm = np.zeros((ny, nx), np.float32); lsm[100:140, 200:260] = 1.0
    coast = oracle.py_get_edges(lsm, np.zeros_like(lsm))
    assert oracle.py_dist_radius(lon, lat) == 2                      # k = 2 on ERA-Interim
    cd = oracle.py_get_dist(coast, lsm, lon, lat)
    assert (np.abs(cd[coast > 0]) == 0.5).all()                      # B.3 self distance exactly 0.5
    assert cd[10, 10] == 12000.0                                     # nothing within the window
    assert (cd[(lsm == 0) & (np.abs(cd) < 12000)] < 0).all()         # sea negative
    assert (cd[lsm > 0] > 0).all()


def _uniform_case(nx=24, ny=16, nps=5):
    lsm = np.zeros((ny, nx), np.float32); lsm[:, 12:] = 1.0
    cd = np.where(lsm > 0, 10.0, -10.0).astype(np.float32)
    theta = np.where(lsm > 0, 300.0, 298.0).astype(np.float32)
    z = np.zeros_like(lsm); std = np.zeros_like(lsm); std[0, 0] = 1.0   # non-constant sigma (A.6)
    p = np.array([100000, 85000, 70000, 50000, 30000], np.float32) / 1.0
    p_hpa_target = 700.0
    u = np.zeros((nps, ny, nx), np.float32); v = np.zeros_like(u)
    return lsm, cd, theta, z, std, p, u, v, p_hpa_target


def test_diag_first_step_known_answer(oracle):
    """B.4: land 300 K / sea 298 K, |V| = 5 m/s => sb_con = +-0.75 next to the coast."""
    lsm, cd, theta, z, std, p, u, v, tp = _uniform_case()
    u[2] = 3.0; v[2] = 4.0                                           # 700 hPa is level index 2
    out, fl = oracle.py_diag(1, p, z, std, theta, v, u, cd, np.zeros_like(z), np.zeros_like(z))
    assert out[0][5, 12] == np.float32(0.75)                         # land cell next to the coast
    assert out[0][5, 11] == np.float32(-0.75)                        # adjacent sea cell
    assert fl[5, 12] == 3 and fl[5, 11] == 3
    assert out[2][5, 12] == 5.0                                      # windspeed
    assert (out[0][-1] == np.float32(2.0e20)).all()                  # last row never processed (P:165)


def test_diag_winddir_known_answer(oracle):
    """B.5: u=0,v=-1 -> 0 deg ; u=-1,v=0 -> atan2(1,0)*57.2957 = 89.99988 (truncated rad2deg)."""
    lsm, cd, theta, z, std, p, u, v, tp = _uniform_case()
    u[2] = 0.0; v[2] = -1.0
    out, _ = oracle.py_diag(1, p, z, std, theta, v, u, cd, np.zeros_like(z), np.zeros_like(z))
    assert out[3][5, 12] == 0.0
    u[2] = -1.0; v[2] = 0.0
    for mode in (0, 1):
        oracle.set_math(mode)
        out, _ = oracle.py_diag(1, p, z, std, theta, v, u, cd, np.zeros_like(z), np.zeros_like(z))
        assert abs(float(out[3][5, 12]) - 89.99988) < 1e-5
    oracle.set_math(1)


def test_diag_thresholds(oracle):
    """B.7: ws = 12 => no trigger; |n_thc| == 0.75 exactly => no trigger (strict >)."""
    lsm, cd, theta, z, std, p, u, v, tp = _uniform_case()
    u[2] = 12.0
    out, fl = oracle.py_diag(1, p, z, std, theta, v, u, cd, np.zeros_like(z), np.zeros_like(z))
    assert (fl & 2).sum() == 0 and (out[0][:-1][fl[:-1] == 1] == 0).all()
    u[2] = 3.0; v[2] = 4.0
    theta2 = np.where(lsm > 0, 298.75, 298.0).astype(np.float32)
    out, fl = oracle.py_diag(1, p, z, std, theta2, v, u, cd, np.zeros_like(z), np.zeros_like(z))
    assert (fl & 2).sum() == 0


@pytest.mark.parametrize("timestep_min,expected", [(24.0, [15, 30, 45]), (60.0, [6, 12, 18, 24])])
def test_diag_six_hourly_update(oracle, timestep_min, expected):
    """B.6: the windspeed/winddir state is refreshed only when step*dt is a multiple of 6 h."""
    lsm, cd, theta, z, std, p, u, v, tp = _uniform_case()
    ws = np.zeros_like(z); wd = np.zeros_like(z)
    fired = []
    for step in range(1, 50 if timestep_min == 24.0 else 25):
        u[2] = 1.0 + 0.01 * step; v[2] = 0.5
        out, _ = oracle.py_diag(step, p, z, std, theta, v, u, cd, ws, wd, timestep=timestep_min)
        new_ws = np.float32(np.sqrt(np.float32(u[2, 5, 12] ** 2 + v[2, 5, 12] ** 2)))
        if step > 1 and out[2][5, 12] == new_ws and ws[5, 12] != new_ws:
            fired.append(step)
        ws, wd = out[2].copy(), out[3].copy()
    assert fired == expected


def test_sigmoid_sequential_order(oracle):
    """G:466-478: the statistics are sequential fp32 accumulations, not pairwise/fp64."""
    rng = np.random.default_rng(3)
    a = (rng.random((200, 300)) * 300).astype(np.float32)
    sm, st = oracle.sigmoid(a)
    s = np.float32(0)
    for x in a.ravel():
        s = np.float32(s + x)
    mean = np.float32(s / np.float32(a.size))
    assert st[0] == mean
    var = np.float32(0)
    for x in a.ravel():
        d = np.float32(x - mean); var = np.float32(var + np.float32(d * d))
    assert st[3] == var
    assert st[1] == np.float32(np.float32(2.0) / np.sqrt(np.float32(var / np.float32(a.size))))
    assert st[2] == np.float32((a.max() - a.min()) / np.float32(4.0))
    assert 0 < sm.min() and sm.max() <= 1


def test_byproducts_known_answers(oracle):
    """orc_sobel / orc_window_info on hand-checkable cases."""
    h = 2
    # meridional coastline: land for x >= 4 in an 8 x 6 grid (halo filled by swap_bounds: E-W periodic)
    m = np.zeros((6, 8), np.float32); m[:, 4:] = 1.0
    mh = np.zeros((6 + 2 * h, 8 + 2 * h), np.float32); mh[h:-h, h:-h] = m
    oracle.swap_bounds(mh, h)
    px, py = oracle.sobel(mh, h)
    assert np.all(py == 0.0)                       # no north-south gradient
    assert np.all(px[:, 3] == 4.0) and np.all(px[:, 4] == 4.0)     # (1+2+1) * (1 - 0) on both coast columns
    assert np.all(px[:, 1:3] == 0.0) and np.all(px[:, 5:7] == 0.0)
    assert np.all(px[:, 0] == -4.0) and np.all(px[:, 7] == -4.0)   # the periodic seam is a coastline too
    # signed distance plane: sea (-) left of x = 4, land (+) right; everything inside maxdist
    cd = np.where(m > 0, 10.0, -10.0).astype(np.float32)
    ch = np.zeros_like(mh); ch[h:-h, h:-h] = cd
    oracle.swap_bounds(ch, h)
    nn, nl, ns = oracle.window_info(ch, h)
    assert np.all(nn[:, 3] == 1) and np.all(nn[:, 4] == 1)         # next to the coast: 3 x 3 window suffices
    assert np.all(nl[:, 3] == 3) and np.all(ns[:, 3] == 6) and np.all(nl[:, 4] == 6) and np.all(ns[:, 4] == 3)
    assert np.all(nn[:, 2] == 2) and np.all(nl[:, 2] == 5) and np.all(ns[:, 2] == 20)   # 5 x 5: one land column
    assert np.all((nl + ns) == (2 * nn + 1) ** 2)
    # cells beyond maxdist are skipped (G:174-176)
    ch2 = ch.copy(); ch2[h:-h, h:h + 2] = -500.0
    nn2, nl2, ns2 = oracle.window_info(ch2, h)
    assert np.all(nn2[:, :2] == 0) and np.all(nl2[:, :2] == 0) and np.all(ns2[:, :2] == 0)
