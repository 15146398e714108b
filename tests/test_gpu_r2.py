"""Round-2 parity cases ("What's weak: parity" 2-5): search radii beyond 15 cells for get_dist,
the static chain at c2 full size, and the reference's OWN python wrapper running on the drop-in module.
All comparisons are bit-exact against the oracle through the C-ABI.  Runs on the B200 box: -m gpu."""
import os
import subprocess
import sys

import numpy as np
import pytest

from seabreeze_param_b200 import generic, seabreeze, seabreezediag, synth
from util import assert_bits_equal

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def F(a):
    return seabreezediag.c2f(np.ascontiguousarray(a))


def _regional_block(nx, ny, filt, seed=3):
    """A block with the c3 grid spacing (0.0225 degree, descending latitude): get_dist derives k = 54."""
    cfg = synth.Config("blk", nx, ny, 2, 54, "generic", filt, True, 1440.0, "c3-spaced block")
    return cfg, synth.static_fields(cfg, seed=seed)


@pytest.mark.parametrize("k", [24, 54])
def test_generic_get_dist_wide_radius_vs_literal_scatter(oracle, k):
    """generic/sea_breeze_diag.f90:419-453 with halo_size = 24 and 54 (c3's radius): the GPU gather
    against the oracle's literal scatter (sobel.f90:152-192 order, visit-time reset S:188) and against
    the oracle's own gather form."""
    cfg, (land, ice, lon, lat, z, sig) = _regional_block(640, 3 * k + 40, 12.0)
    coast_h, _ = oracle.get_edges(ice, land, k)
    assert coast_h.sum() > 500
    cd_lit = oracle.get_dist(coast_h, land, lon, lat, 180.0, k, literal=True)
    oracle.set_threads(oracle.max_threads())
    try:
        cd_gat = oracle.get_dist(coast_h, land, lon, lat, 180.0, k, literal=False)
    finally:
        oracle.set_threads(1)
    assert_bits_equal(cd_lit, cd_gat, "oracle scatter vs gather, k=%d" % k)
    mask_g = np.zeros_like(coast_h)
    generic.get_edges(mask_g, ice, land, k)
    assert_bits_equal(coast_h, mask_g, "coast k=%d" % k)
    cd_g = np.empty_like(land)
    generic.get_dist(mask_g, land, lon, lat, 180.0, cd_g, k)
    assert_bits_equal(cd_lit, cd_g, "get_dist k=%d" % k)
    assert (np.abs(cd_g) < 360.0).any()


def test_py_get_dist_radius_50_vs_literal_scatter(oracle):
    """python_wrapper/seabreezediag/sobel.f90:129-137 derives the radius itself: 51 cells on this block of the
    2.5 km grid (its first row stands in for the 70-degree row), 22 with maxdist = 80."""
    cfg, (land, ice, lon, lat, z, sig) = _regional_block(700, 230, 12.0, seed=5)
    coast = oracle.py_get_edges(land, ice)
    for maxdist, kmin in ((180.0, 48), (80.0, 20)):
        assert oracle.py_dist_radius(lon, lat, maxdist) >= kmin
        cd_o = oracle.py_get_dist(coast, land, lon, lat, maxdist=maxdist, literal=True)
        cd_g = seabreeze.get_dist(F(coast), F(land), lon, lat, maxdist=maxdist)
        assert_bits_equal(F(cd_o), cd_g, "python get_dist maxdist=%g" % maxdist)


def test_c2_static_chain_full_size(oracle):
    """get_edges -> get_dist at the benchmark grid (1024x769, halo 6) against the oracle's literal loops;
    the search radius / land-count by-products of the plan against orc_window_info."""
    cfg = synth.CONFIGS["c2"]
    h = cfg.h
    land, ice, lon, lat, z, sig = synth.static_fields(cfg)
    coast_o, m1_o = oracle.get_edges(ice, land, h)
    cd_o = oracle.get_dist(coast_o, land, lon, lat, 180.0, h, literal=True)
    mask_g = np.zeros_like(coast_o)
    m1_g = np.zeros((cfg.ny, cfg.nx), np.uint8)
    generic.get_edges(mask_g, ice, land, h, m1=m1_g)
    assert_bits_equal(coast_o, mask_g, "c2 coast")
    assert np.array_equal(m1_o, m1_g)
    cd_g = np.empty_like(land)
    generic.get_dist(mask_g, land, lon, lat, 180.0, cd_g, h)
    assert_bits_equal(cd_o, cd_g, "c2 cdist")
    mk = generic.swap_bounds(synth.pad_halo(cd_g, h), h)
    zh = generic.swap_bounds(synth.pad_halo(z, h), h)
    sh = generic.swap_bounds(synth.pad_halo(sig, h), h)
    plan = generic.Plan(generic.SBD_VARIANT_GENERIC, cfg.nx, cfg.ny, cfg.nz, h)
    plan.set_static(mk, zh, sh)
    nn_g, nl_g, ns_g = plan.window_info()
    oracle.set_threads(oracle.max_threads())
    try:
        nn_o, nl_o, ns_o = oracle.window_info(mk, h)
    finally:
        oracle.set_threads(1)
    assert np.array_equal(nn_o, nn_g) and np.array_equal(nl_o, nl_g) and np.array_equal(ns_o, ns_g)
    _, st_o = oracle.sigmoid(sig)
    info = plan.info()
    assert_bits_equal(st_o[:3], np.array([info["mean"], info["std2inv"], info["r"]], np.float32), "c2 sigma statistics")
    plan.close()


def _reference_wrapper_dir():
    for d in ("/root/reference/python_wrapper", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.exists(os.path.join(d, "seabreezediag", "__init__.py")):
            return d
    return None


def test_reference_python_wrapper_runs_on_the_dropin(oracle, tmp_path):
    """BASELINE.json configs[4]: the reference's own python_wrapper/seabreezediag/__init__.py (unmodified:
    from /root/reference, or the copy scripts/install_reference.py leaves in the git-ignored baseline/_ref)
    imported with our module named `seabreeze` on its path; its diag() (I:91-263) is run on the ERA-Interim-size
    grid (480x241, 37 levels, python_wrapper/run.conf:5) for three 6-hourly records with time-varying sea ice
    and compared with the oracle driven through the same call sequence (I:219-245)."""
    ref = _reference_wrapper_dir()
    if ref is None:
        pytest.skip("the reference's python wrapper is neither under /root/reference nor under baseline/_ref")
    cfg = synth.CONFIGS["c5"]
    land, ice, lon, lat, z, sig = synth.static_fields(cfg)
    nrec = 3
    th, uu, vv, ci = [], [], [], []
    for r in range(nrec):
        theta, p, u, v = synth.step_fields(cfg, r + 1, land, z, lat)
        th.append(theta); uu.append(u); vv.append(v)
        ci.append(np.clip(ice * (1.0 + 0.5 * r), 0, 1).astype(np.float32))
    th, uu, vv, ci = (np.stack(x) for x in (th, uu, vv, ci))
    inp = str(tmp_path / "in.npz")
    outp = str(tmp_path / "out.npz")
    np.savez(inp, land=land, z=z, sig=sig, lon=lon, lat=lat, p=p, th=th, uu=uu, vv=vv, ci=ci)
    code = (
        "import sys, numpy as np\n"
        "sys.path[:0] = [%r, %r, %r]\n"
        "import seabreezediag as sbd, seabreeze\n"
        "assert sbd.__file__.startswith(%r), sbd.__file__\n"
        "assert sbd.dg is seabreeze.diag\n"
        "d = np.load(%r)\n"
        "tt, sb, thc, ws, wd = sbd.diag(1, d['land'], d['z'], d['sig'], d['lon'], d['lat'], d['p'],\n"
        "                               d['uu'], d['vv'], d['th'], d['ci'], timestep=%r)\n"
        "np.savez(%r, tt=tt, sb=sb, thc=thc, ws=ws, wd=wd)\n"
        % (os.path.join(ROOT, "seabreeze_param_b200", "dropin"), ROOT, ref, ref, inp, cfg.timestep, outp))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    got = np.load(outp)
    assert int(got["tt"]) == 1 + nrec
    # oracle, same call sequence
    ws = np.zeros_like(land); wd = np.zeros_like(land)
    trig = 0
    for rec in range(nrec):
        coast = oracle.py_get_edges(land, ci[rec])
        cd = oracle.py_get_dist(coast, land, lon, lat, literal=True)
        o, f = oracle.py_diag(1 + rec, p, z, sig, th[rec], vv[rec], uu[rec], cd, ws, wd, timestep=cfg.timestep,
                              literal=True)
        interior = slice(0, cfg.ny - 1)   # the last row is never written by the reference (P:165)
        assert_bits_equal(o[0][interior], got["sb"][rec][interior].astype(np.float32), "sb_con record %d" % rec)
        ws, wd = o[2].copy(), o[3].copy()
        trig += int((f & 2).sum())
        last_t0 = o[1]
    assert trig > 0
    assert_bits_equal(ws, got["ws"].astype(np.float32), "windspeed")
    assert_bits_equal(wd, got["wd"].astype(np.float32), "winddir")
    assert_bits_equal(last_t0, got["thc"].astype(np.float32), "thc (= t0, I:244)")


@pytest.mark.parametrize("name", ["ref_py_a.npz", "ref_py_b.npz"])
def test_cuda_path_reproduces_reference_source_goldens(name):
    """tests/golden/ref_*.npz were produced by the reference's own .f90 files (executed by tests/golden/f90interp.py,
    see make_reference_golden.py): the CUDA path, through the f2py-compatible module, must give the same bits --
    get_edges, get_dist and the diag() records with the state threaded as the reference wrapper does."""
    d = np.load(os.path.join(ROOT, "tests", "golden", name))
    land, z, sig, lon, lat, p = (d[k] for k in ("land", "z", "sig", "lon", "lat", "p"))
    kw = {k[3:]: float(d[k]) for k in d.files if k.startswith("kw_")}
    ws = np.zeros_like(z); wd = np.zeros_like(z); thc = np.zeros_like(z)
    for r in range(d["theta"].shape[0]):
        coast = seabreeze.get_edges(F(land), F(d["ci"][r]))
        assert_bits_equal(F(d["coast"][r]), coast, "%s get_edges record %d" % (name, r))
        cd = seabreeze.get_dist(coast, F(land), lon, lat, maxdist=float(d["dist_maxdist"]))
        assert_bits_equal(F(d["cdist"][r]), cd, "%s get_dist record %d" % (name, r))
        out = seabreeze.diag(1 + r, p, F(z), F(sig), F(d["theta"][r]), F(d["v"][r]), F(d["u"][r]), cd, F(ws), F(wd),
                             F(thc), **kw)
        got = seabreezediag.c2f(out)
        ref = d["out"][r]
        for k, nm in enumerate(("sb_con", "t0", "windspeed", "winddir")):
            w = ~np.isnan(ref[k])     # NaN: never written by the reference (last row, P:165)
            assert_bits_equal(np.where(w, ref[k], 0).astype(np.float32), np.where(w, got[k], 0).astype(np.float32),
                              "%s %s record %d" % (name, nm, r))
        thc, ws, wd = [np.where(np.isnan(ref[k]), prev, ref[k]).astype(np.float32)
                       for k, prev in ((1, thc), (2, ws), (3, wd))]


def test_one_shot_plan_cache_and_host_registration(oracle):
    """sbd_seabreeze_diag_f32 -- the call dummy_model.f90:52-54 makes every timestep -- keeps its plan between
    calls: same static fields -> cache hit (no rebuild), changed static field -> rebuild, and the results are the
    oracle's in both cases.  Large host arrays handed to generic.seabreeze_diag get page-locked (auto_register)."""
    import ctypes as C
    from seabreeze_param_b200 import _capi
    from util import generic_case, generic_step_inputs
    lib = _capi.load()
    lib.sbd_plan_cache_clear()
    cfg = synth.Config("t_cache", 512, 384, 24, 5, "generic", 6.0, False, 1440.0, "plan cache test")
    case = generic_case(cfg, oracle)
    ny, nx = cfg.ny, cfg.nx
    so = [np.zeros((ny, nx), np.float32) for _ in range(4)]
    sg = [np.zeros((ny, nx), np.float32) for _ in range(4)]
    h0, m0, n0 = C.c_int(), C.c_int(), C.c_int()
    lib.sbd_plan_cache_stats(C.byref(h0), C.byref(m0), C.byref(n0))

    def step(k, mask_h):
        th, p, u, v = generic_step_inputs(cfg, case, k, oracle)
        fo = oracle.seabreeze_diag(cfg.timestep, k, p, u, v, th, mask_h, case["z_h"], case["sigma_h"], *so, cfg.h)
        fg = np.zeros((ny, nx), np.uint8)
        generic.seabreeze_diag(cfg.timestep, k, p, u, v, th, mask_h, case["z_h"], case["sigma_h"], *sg, flags=fg)
        for a, b, nm in zip(so, sg, ("windspeed", "winddir", "thc", "sb_con")):
            assert_bits_equal(a, b, "%s step %d" % (nm, k))
        assert np.array_equal(fo, fg)
        return p

    p = step(1, case["mask_h"])
    step(2, case["mask_h"])
    step(3, case["mask_h"])
    h1, m1, n1 = C.c_int(), C.c_int(), C.c_int()
    lib.sbd_plan_cache_stats(C.byref(h1), C.byref(m1), C.byref(n1))
    assert m1.value - m0.value == 1 and h1.value - h0.value == 2 and n1.value >= 1
    # a changed static field must be noticed: move the coast (sign flip of one patch of the signed distance)
    mask2 = case["mask_h"].copy()
    mask2[100:140, 200:260] *= -1.0
    step(4, mask2)
    lib.sbd_plan_cache_stats(C.byref(h1), C.byref(m1), C.byref(n1))
    assert m1.value - m0.value == 2
    assert p.nbytes >= (1 << 20) and p.ctypes.data in _capi._REGISTERED    # 18.9 MB p was page-locked
    lib.sbd_plan_cache_clear()


def test_sequential_statistics_cross_check():
    """sbd_set_stats_mode(1): the plain one-warp sequential chain (k_stats_seq) instead of the exact parallel
    formulation (k_stats_exact) -- the same bits, on a field with ties, zeros and binade crossings."""
    from seabreeze_param_b200 import _capi
    lib = _capi.load()
    rng = np.random.default_rng(3)
    a = (rng.random((300, 517)) * 300 * (rng.random((300, 517)) > 0.6)).astype(np.float32)
    a[7, :40] = 2.5
    st_p = np.zeros(4, np.float32); st_s = np.zeros(4, np.float32)
    sm_p = generic.sigmoid(a, stats=st_p)
    lib.sbd_set_stats_mode(1)
    try:
        sm_s = generic.sigmoid(a, stats=st_s)
    finally:
        lib.sbd_set_stats_mode(0)
    assert_bits_equal(st_p, st_s, "statistics")
    assert_bits_equal(sm_p, sm_s, "sigmoid")
