"""The oracle's literal evaluation (the reference's loops as written) must agree bit-for-bit with
its fast-but-equivalent evaluation (gather get_dist, final-window sums), and the pinned math must
stay within tolerance of the platform libm (what gfortran would call)."""
import numpy as np
import pytest

from seabreeze_param_b200 import synth
from util import assert_bits_equal, generic_case, generic_step_inputs


@pytest.mark.parametrize("name", ["t_py", "c5"])
def test_python_variant_literal_equals_fast(oracle, name):
    cfg = synth.CONFIGS[name]
    land, ice, lon, lat, z, sig = synth.static_fields(cfg)
    coast = oracle.py_get_edges(land, ice)
    cd_l = oracle.py_get_dist(coast, land, lon, lat, literal=True)
    cd_g = oracle.py_get_dist(coast, land, lon, lat, literal=False)
    assert_bits_equal(cd_l, cd_g, "get_dist scatter vs gather")
    ws = np.zeros_like(z); wd = np.zeros_like(z)
    for step in (1, 2, 3):
        theta, p, u, v = synth.step_fields(cfg, step, land, z, lat)
        o1, f1 = oracle.py_diag(step, p, z, sig, theta, v, u, cd_l, ws, wd, timestep=cfg.timestep, literal=True)
        o2, f2 = oracle.py_diag(step, p, z, sig, theta, v, u, cd_l, ws, wd, timestep=cfg.timestep, literal=False)
        assert_bits_equal(o1, o2, "diag literal vs fast, step %d" % step)
        assert np.array_equal(f1, f2)
        ws, wd = o1[2].copy(), o1[3].copy()
    assert (f1 & 2).sum() > 0, "synthetic case never triggers: test is vacuous"


def test_dist_reset_rule_matters(oracle):
    """S:188 resets |cdist| > 2*maxdist at visit time; with a wide window the gather form must
    reproduce the earlier/later asymmetry."""
    cfg = synth.CONFIGS["t_py"]
    land, ice, lon, lat, z, sig = synth.static_fields(cfg)
    coast = oracle.py_get_edges(land, ice)
    for k in (3, 5):
        cd_l = oracle.py_get_dist(coast, land, lon, lat, maxdist=180.0, k=k, literal=True)
        cd_g = oracle.py_get_dist(coast, land, lon, lat, maxdist=180.0, k=k, literal=False)
        assert_bits_equal(cd_l, cd_g, "k=%d" % k)
        assert (np.abs(cd_l) > 360.0).any() and (np.abs(cd_l)[np.abs(cd_l) < 12000] > 360.0).any()


@pytest.mark.parametrize("level_scan", [0, 1])
def test_generic_variant_literal_equals_fast(oracle, level_scan):
    cfg = synth.CONFIGS["t_gen"]
    case = generic_case(cfg, oracle)
    cd_lit = oracle.get_dist(case["coast_h"], case["land"], case["lon"], case["lat"], 180.0, cfg.h, literal=True)
    assert_bits_equal(cd_lit, case["cdist"], "generic get_dist scatter vs gather")
    ny, nx = cfg.ny, cfg.nx
    st = [[np.zeros((ny, nx), np.float32) for _ in range(4)] for _ in range(2)]
    for step in (1, 2, 15):
        theta_h, p, u, v = generic_step_inputs(cfg, case, step, oracle)
        fl = []
        for lit in (0, 1):
            ws, wd, thc, sb = st[lit]
            fl.append(oracle.seabreeze_diag(cfg.timestep, step, p, u, v, theta_h, case["mask_h"], case["z_h"],
                                            case["sigma_h"], ws, wd, thc, sb, cfg.h, literal=bool(lit),
                                            level_scan=level_scan))
        for a, b, nm in zip(st[0], st[1], ("ws", "wd", "thc", "sb")):
            assert_bits_equal(a, b, "%s step %d" % (nm, step))
        assert np.array_equal(fl[0], fl[1])
    assert (fl[0] & 2).sum() > 0


def test_pinned_vs_libm_tolerance(oracle):
    """Pinned transcendental algorithms vs the platform libm on a full python-variant record:
    flags identical, floats within rtol 1e-5 / atol 1e-4 (BASELINE.md parity bar)."""
    cfg = synth.CONFIGS["c5"]
    land, ice, lon, lat, z, sig = synth.static_fields(cfg)
    res = {}
    for mode in (0, 1):
        oracle.set_math(mode)
        coast = oracle.py_get_edges(land, ice)
        cd = oracle.py_get_dist(coast, land, lon, lat, literal=False)
        theta, p, u, v = synth.step_fields(cfg, 1, land, z, lat)
        o, f = oracle.py_diag(1, p, z, sig, theta, v, u, cd, np.zeros_like(z), np.zeros_like(z),
                              timestep=cfg.timestep, literal=False)
        res[mode] = (cd, o, f)
    oracle.set_math(1)
    assert np.array_equal(res[0][2], res[1][2]), "flag mismatches between libm and pinned math"
    np.testing.assert_allclose(res[0][0], res[1][0], rtol=1e-5, atol=1e-4)
    good = res[0][2] > 0
    for k in range(4):
        np.testing.assert_allclose(res[0][1][k][good], res[1][1][k][good], rtol=1e-5, atol=1e-4)
