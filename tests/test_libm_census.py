"""Pinned-math vs platform-libm census on the benchmark grids (BASELINE.md parity bar).

The oracle evaluates every transcendental either with the pinned algorithms of include/sbd_math.h (what
the GPU runs; mode 1) or with the platform libm that a gfortran build of the reference would call
(expf/sinf/cosf/atan2f; mode 0).  On c2 (348 k coastal cells x 3 steps) and c4 (1.97 M coastal cells) the
two modes are compared over every output: the test prints the number of flag mismatches and of
floats whose bits differ, requires identical flags and rtol 1e-5 / atol 1e-4 on the floats (the
libm-mode tolerance that stands next to the bit-exact pinned-math claim).  Measured in this
container (glibc 2.39): 0 differing bits anywhere -- recorded in DESIGN.md section 2.  CPU only."""
import numpy as np

from seabreeze_param_b200 import synth
from util import generic_case, generic_step_inputs


def _report(tag, a, b):
    nb = int((np.ascontiguousarray(a).view(np.uint32) != np.ascontiguousarray(b).view(np.uint32)).sum())
    print("census %-28s differing bit patterns: %d of %d" % (tag, nb, a.size))
    return nb


def test_census_c2_generic(oracle):
    cfg = synth.CONFIGS["c2"]
    oracle.set_threads(oracle.max_threads())
    res = {}
    try:
        for mode in (0, 1):
            oracle.set_math(mode)
            case = generic_case(cfg, oracle)
            st = [np.zeros((cfg.ny, cfg.nx), np.float32) for _ in range(4)]
            out = []
            for s in (1, 2, 15):   # first step, ordinary step, 6-hourly step
                th, p, u, v = generic_step_inputs(cfg, case, s, oracle)
                f = oracle.seabreeze_diag(cfg.timestep, s, p, u, v, th, case["mask_h"], case["z_h"], case["sigma_h"],
                                          *st, cfg.h, literal=False)
                out.append((f.copy(), [a.copy() for a in st]))
            res[mode] = (case["cdist"].copy(), out)
    finally:
        oracle.set_math(1)
        oracle.set_threads(1)
    _report("c2 cdist", res[0][0], res[1][0])
    zone0, zone1 = np.abs(res[0][0]) <= 180.0, np.abs(res[1][0]) <= 180.0
    assert np.array_equal(zone0, zone1), "coastal-zone mask (m3) differs between libm and pinned math"
    np.testing.assert_allclose(res[0][0], res[1][0], rtol=1e-5, atol=1e-4)
    for i, s in enumerate((1, 2, 15)):
        f0, f1 = res[0][1][i][0], res[1][1][i][0]
        print("census c2 step %2d: coastal %d, trigger %d, flag mismatches %d" % (
            s, int((f1 & 1).sum()), int(((f1 & 2) > 0).sum()), int((f0 != f1).sum())))
        assert np.array_equal(f0, f1), "c2 step %d: trigger flags differ between libm and pinned math" % s
        for k, nm in enumerate(("windspeed", "winddir", "thc", "sb_con")):
            a, b = res[0][1][i][1][k], res[1][1][i][1][k]
            _report("c2 step %d %s" % (s, nm), a, b)
            np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-4, equal_nan=True)


def test_census_c4_python(oracle):
    cfg = synth.CONFIGS["c4"]
    land, ice, lon, lat, z, sig = synth.static_fields(cfg)
    theta, p, u, v = synth.step_fields(cfg, 1, land, z, lat)
    oracle.set_threads(oracle.max_threads())
    res = {}
    try:
        for mode in (0, 1):
            oracle.set_math(mode)
            coast = oracle.py_get_edges(land, ice)
            cd = oracle.py_get_dist(coast, land, lon, lat, literal=False)
            o, f = oracle.py_diag(1, p, z, sig, theta, v, u, cd, np.zeros_like(z), np.zeros_like(z),
                                  timestep=cfg.timestep, literal=False)
            res[mode] = (cd, o, f)
    finally:
        oracle.set_math(1)
        oracle.set_threads(1)
    _report("c4 cdist", res[0][0], res[1][0])
    np.testing.assert_allclose(res[0][0], res[1][0], rtol=1e-5, atol=1e-4)
    f0, f1 = res[0][2], res[1][2]
    print("census c4 record 1: coastal %d, trigger %d, flag mismatches %d" % (
        int((f1 & 1).sum()), int(((f1 & 2) > 0).sum()), int((f0 != f1).sum())))
    assert np.array_equal(f0, f1), "c4: trigger flags differ between libm and pinned math"
    good = f1 > 0
    for k, nm in enumerate(("sb_con", "t0", "windspeed", "winddir")):
        _report("c4 " + nm, res[0][1][k], res[1][1][k])
        np.testing.assert_allclose(res[0][1][k][good], res[1][1][k][good], rtol=1e-5, atol=1e-4)
