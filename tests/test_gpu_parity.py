"""Parity of the CUDA path (through the C-ABI) against the CPU oracle, bit-exact: masks, flags and
every floating-point diagnostic (both sides run the pinned transcendental algorithms of
include/sbd_math.h with identical IEEE operation order).  Runs on the B200 box: -m gpu."""
import numpy as np
import pytest

from seabreeze_param_b200 import generic, seabreeze, seabreezediag, synth
from seabreeze_param_b200._capi import SBD_VARIANT_GENERIC, SBD_VARIANT_PYTHON, load
from util import window_kernel, assert_bits_equal, generic_case, generic_step_inputs

pytestmark = pytest.mark.gpu


def F(a):
    """C-ordered (ny,nx[,..]) oracle array -> the Fortran-ordered (nx,ny) view f2py callers hold."""
    return seabreezediag.c2f(np.ascontiguousarray(a))


def test_device_present_and_library_loaded():
    lib = load()
    assert lib.sbd_device_count() >= 1, "no CUDA device: the product has no CPU fallback"


@pytest.mark.parametrize("name,frac", [("t_py", False), ("c5", False), ("c5", True)])
def test_py_edges_and_dist(oracle, name, frac):
    cfg = synth.CONFIGS[name]
    land, ice, lon, lat, z, sig = synth.static_fields(cfg, fractional=frac)
    coast_o = oracle.py_get_edges(land, ice)
    coast_g = seabreeze.get_edges(F(land), F(ice))
    assert coast_g.flags.f_contiguous and coast_g.dtype == np.float32
    assert_bits_equal(F(coast_o), coast_g, "get_edges")
    cd_o = oracle.py_get_dist(coast_o, land, lon, lat, literal=True)
    cd_g = seabreeze.get_dist(coast_g, F(land), lon, lat)
    assert_bits_equal(F(cd_o), cd_g, "get_dist")


def test_py_edges_wrap_quirk_and_ragged(oracle):
    rng = np.random.default_rng(5)
    for ny, nx in ((3, 2), (5, 33), (17, 31), (2, 64)):
        lsm = (rng.random((ny, nx)) > 0.6).astype(np.float32)
        ci = (rng.random((ny, nx)) * 0.5).astype(np.float32)
        assert_bits_equal(F(oracle.py_get_edges(lsm, ci)), seabreeze.get_edges(F(lsm), F(ci)), "%dx%d" % (nx, ny))


def test_py_dist_wide_window_reset_rule(oracle):
    """Windows wider than 2*maxdist exercise the visit-order reset (S:188)."""
    cfg = synth.CONFIGS["t_py"]
    land, ice, lon, lat, z, sig = synth.static_fields(cfg)
    coast = oracle.py_get_edges(land, ice)
    # the python entry point derives k itself; use the generic one to force k (= halo)
    for k in (3, 5):
        cd_o = oracle.py_get_dist(coast, land, lon, lat, maxdist=180.0, k=k, literal=True)
        coast_h = synth.pad_halo(coast, k)
        cd_g = np.empty_like(land)
        generic.get_dist(coast_h, land, lon, lat, 180.0, cd_g, k)
        assert_bits_equal(cd_o, cd_g, "k=%d" % k)


def test_sigmoid_sequential_stats(oracle):
    rng = np.random.default_rng(11)
    for ny, nx in ((1, 1 + 0), (7, 33), (200, 300), (241, 480)):
        a = (rng.random((ny, nx)) * 300).astype(np.float32)
        if a.size > 1:
            a.ravel()[:: 3] = 0.0
        sm_o, st_o = oracle.sigmoid(a)
        st_g = np.zeros(4, np.float32)
        sm_g = generic.sigmoid(a, stats=st_g)
        assert_bits_equal(st_o, st_g, "stats %dx%d" % (nx, ny))
        assert_bits_equal(sm_o, sm_g, "sigmoid %dx%d" % (nx, ny))
        assert_bits_equal(F(sm_o), seabreeze.sigmoid(F(a)), "f2py sigmoid")


def test_exact_parallel_stats_adversarial(oracle):
    """k_stats_exact (parallel, monoid scan) must reproduce the sequential fp32 chain bit-for-bit on
    inputs built to hit ties (round-half-even), binade crossings, zeros, denormals, huge jumps,
    negatives and non-finite values."""
    rng = np.random.default_rng(21)
    cases = {}
    cases["integers (ties everywhere)"] = rng.integers(0, 5, (300, 301)).astype(np.float32)
    cases["halves"] = (rng.integers(0, 9, (257, 129)) * 0.5).astype(np.float32)
    cases["sparse land (70% zeros)"] = (rng.random((400, 513)) * 300 * (rng.random((400, 513)) > 0.7)).astype(np.float32)
    a = rng.random((97, 1031)).astype(np.float32); a[0, :50] = 0; a[40, 7] = 3.0e7; a[41, 9] = 1.0e-30
    cases["jumps + tiny"] = a
    cases["denormals"] = (rng.random((33, 65)) * 1e-41).astype(np.float32)
    b = (rng.random((64, 200)) * 10).astype(np.float32); b[10, 10] = -3.5; b[50, 3] = -0.0
    cases["negatives"] = b
    c = (rng.random((40, 50)) * 10).astype(np.float32); c[20, 20] = np.inf
    cases["inf"] = c
    cases["powers of two"] = (2.0 ** rng.integers(-3, 12, (150, 150))).astype(np.float32)
    cases["one cell"] = np.array([[7.25]], np.float32)
    # 6.48 M cells: the running sum parks on the last value of a binade (regression: n0 > limit)
    cases["0.1 degree grid size"] = (rng.random((1800, 3600)) * 300 * (rng.random((1800, 3600)) > 0.7)).astype(np.float32)
    for nm, arr in cases.items():
        sm_o, st_o = oracle.sigmoid(arr)
        st_g = np.zeros(4, np.float32)
        sm_g = generic.sigmoid(arr, stats=st_g)
        assert_bits_equal(st_o, st_g, "stats: " + nm)
        assert_bits_equal(sm_o, sm_g, "sigmoid: " + nm)


@pytest.mark.parametrize("name,wk", [("t_py", 0), ("c5", 0), ("t_py", 2), ("c5", 2), ("t_py", 3), ("c5", 3),
                                     ("t_py", 1), ("c5", 1)])
def test_py_diag_time_loop(oracle, name, wk):
    """wk = 0 / 1: the tile kernel k_trigger (the automatic choice on these grids); 2 / 3 force the strip window
    kernels (k_win4 / k_win8)."""
    with window_kernel(wk):
        _py_diag_time_loop(oracle, name)


def _py_diag_time_loop(oracle, name):
    cfg = synth.CONFIGS[name]
    land, ice, lon, lat, z, sig = synth.static_fields(cfg)
    coast = oracle.py_get_edges(land, ice)
    cd = oracle.py_get_dist(coast, land, lon, lat, literal=False)
    ws_o = np.zeros_like(z); wd_o = np.zeros_like(z)
    ws_g = np.zeros_like(z); wd_g = np.zeros_like(z)
    trig = 0
    for step in (1, 2, 3):
        theta, p, u, v = synth.step_fields(cfg, step, land, z, lat)
        o, f = oracle.py_diag(step, p, z, sig, theta, v, u, cd, ws_o, wd_o, timestep=cfg.timestep, literal=True)
        fl = np.zeros((cfg.nx, cfg.ny), np.uint8, order="F")
        g = seabreeze.diag(step, p, F(z), F(sig), F(theta), F(v), F(u), F(cd), F(ws_g), F(wd_g), F(np.zeros_like(z)),
                           timestep=cfg.timestep, flags=fl)
        assert g.shape == (cfg.nx, cfg.ny, 4) and g.flags.f_contiguous
        gc = seabreezediag.c2f(g)  # (4, ny, nx)
        for k, nm in enumerate(("sb_con", "t0", "windspeed", "winddir")):
            assert_bits_equal(o[k], gc[k], "%s step %d" % (nm, step))
        assert np.array_equal(f, seabreezediag.c2f(fl)), "flags step %d" % step
        ws_o, wd_o = o[2].copy(), o[3].copy()
        ws_g, wd_g = np.ascontiguousarray(gc[2]), np.ascontiguousarray(gc[3])
        trig += int((f & 2).sum())
    assert trig > 0


@pytest.mark.parametrize("wk", [0, 1, 2, 3])
def test_py_diag_mid_radius(oracle, wk):
    """Python variant on a 0.5-degree grid: get_dist radius k = 9, search radii up to ~10 cells, so
    the automatic choice (wk = 0) is the 8-cell strip kernel k_win8 (whose 4-strip and single-cell paths
    carry these radii); wk = 1 is the tile kernel, 2 the round-1 strip kernel k_win4, 3 k_win8 forced.
    Same bits as the oracle's literal growing-window loops in every form."""
    cfg = synth.Config("t_py_mid", 720, 361, 9, 0, "python", 10.0, False, 60.0, "0.5 degree python-variant test")
    land, ice, lon, lat, z, sig = synth.static_fields(cfg)
    coast = oracle.py_get_edges(land, ice)
    cd = oracle.py_get_dist(coast, land, lon, lat, literal=False)
    ws_o = np.zeros_like(z); wd_o = np.zeros_like(z)
    ws_g = np.zeros_like(z); wd_g = np.zeros_like(z)
    with window_kernel(wk):
        for step in (1, 2, 6):
            theta, p, u, v = synth.step_fields(cfg, step, land, z, lat)
            o, f = oracle.py_diag(step, p, z, sig, theta, v, u, cd, ws_o, wd_o, timestep=cfg.timestep, literal=True)
            fl = np.zeros((cfg.nx, cfg.ny), np.uint8, order="F")
            g = seabreeze.diag(step, p, F(z), F(sig), F(theta), F(v), F(u), F(cd), F(ws_g), F(wd_g),
                               F(np.zeros_like(z)), timestep=cfg.timestep, flags=fl)
            gc = seabreezediag.c2f(g)
            for k, nm in enumerate(("sb_con", "t0", "windspeed", "winddir")):
                assert_bits_equal(o[k], gc[k], "%s step %d" % (nm, step))
            assert np.array_equal(f, seabreezediag.c2f(fl)), "flags step %d" % step
            ws_o, wd_o = o[2].copy(), o[3].copy()
            ws_g, wd_g = np.ascontiguousarray(gc[2]), np.ascontiguousarray(gc[3])


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("name", ["t_py", "t_py_mid"])
def test_py_row_blocks_no_exchange(oracle, name, world):
    """The python variant sharded into overlapping row blocks (halo.PyRowBlock, one plan per rank, no
    exchange): the ranks' owned rows, stitched together, equal the oracle's single-domain result bit
    for bit -- all four output planes, three records with the state threaded through.  The ranks run
    one after the other on this GPU; nothing they compute depends on another rank."""
    from seabreeze_param_b200 import halo
    cfg = synth.CONFIGS[name] if name in synth.CONFIGS else synth.Config(
        "t_py_mid", 720, 361, 9, 0, "python", 10.0, False, 60.0, "0.5 degree python-variant test")
    land, ice, lon, lat, z, sig = synth.static_fields(cfg)
    coast = oracle.py_get_edges(land, ice)
    cd = oracle.py_get_dist(coast, land, lon, lat, literal=False)
    ny, nx = cfg.ny, cfg.nx
    probe = generic.Plan(SBD_VARIANT_PYTHON, nx, ny, cfg.nz)
    probe.set_static(cd, z, sig)
    overlap = probe.info()["nn_max"]
    probe.close()
    blocks = [halo.PyRowBlock(nx, ny, cfg.nz, r, world, cd, z, sig, overlap) for r in range(world)]
    assert all(b.required_overlap() <= overlap for b in blocks)
    assert blocks[0].r0 == 0 and blocks[-1].r1 == ny and all(blocks[i].r1 == blocks[i + 1].r0 for i in range(world - 1))
    ws_o = np.zeros_like(z); wd_o = np.zeros_like(z)
    ws_g = np.zeros_like(z); wd_g = np.zeros_like(z)
    for step in (1, 2, 6):
        theta, p, u, v = synth.step_fields(cfg, step, land, z, lat)
        o, f = oracle.py_diag(step, p, z, sig, theta, v, u, cd, ws_o, wd_o, timestep=cfg.timestep, literal=True)
        got = np.zeros((4, ny, nx), np.float32)
        fl = np.zeros((ny, nx), np.uint8)
        for b in blocks:
            a_, b_ = b.rows()
            c = lambda x: np.ascontiguousarray(x[..., a_:b_, :])
            fb = np.zeros((b_ - a_, nx), np.uint8)
            res = b.step(step, cfg.timestep, p, c(u), c(v), c(theta), c(ws_g), c(wd_g), flags=fb)
            for k in range(4):
                got[k, b.r0:b.r1] = res[k]
            fl[b.r0:b.r1] = fb[b.r0 - a_:b.r1 - a_]
        for k, nm in enumerate(("sb_con", "t0", "windspeed", "winddir")):
            assert_bits_equal(o[k], got[k], "%s step %d (%d row blocks)" % (nm, step, world))
        assert np.array_equal(f, fl), "flags step %d" % step
        ws_o, wd_o = o[2].copy(), o[3].copy()
        ws_g, wd_g = got[2].copy(), got[3].copy()
    for b in blocks:
        b.plan.close()


def test_py_wrapper_diag_matches_manual_loop(oracle):
    """seabreezediag.diag (mirror of the reference's Python API) == the same loop done by hand on
    the oracle, including get_edges/get_dist per record with time-varying ice."""
    cfg = synth.CONFIGS["t_py"]
    land, ice, lon, lat, z, sig = synth.static_fields(cfg)
    nrec = 3
    T, U, V, CI = [], [], [], []
    for step in range(1, nrec + 1):
        theta, p, u, v = synth.step_fields(cfg, step, land, z, lat)
        T.append(theta); U.append(u); V.append(v); CI.append(ice * (1.0 if step != 2 else 0.5))
    T, U, V, CI = (np.stack(x).astype(np.float32) for x in (T, U, V, CI))
    tt, sb, thc, ws, wd = seabreezediag.diag(1, land, z, sig, lon, lat, p, U, V, T, CI, timestep=cfg.timestep)
    assert tt == nrec + 1 and sb.shape == (nrec, cfg.ny, cfg.nx)
    ws_o = np.zeros_like(z); wd_o = np.zeros_like(z)
    for r in range(nrec):
        coast = oracle.py_get_edges(land, CI[r])
        cd = oracle.py_get_dist(coast, land, lon, lat)
        o, _ = oracle.py_diag(r + 1, p, z, sig, T[r], V[r], U[r], cd, ws_o, wd_o, timestep=cfg.timestep)
        assert_bits_equal(o[0], sb[r].astype(np.float32), "sb_con record %d" % r)
        ws_o, wd_o = o[2].copy(), o[3].copy()
    assert_bits_equal(ws_o, ws, "ws"); assert_bits_equal(wd_o, wd, "wd"); assert_bits_equal(o[1], thc, "t0")


def test_py_batched_time_loop(oracle):
    """seabreezediag.diag_batched (one C-ABI call for the whole time loop) ==
    the reference's per-record loop on the oracle, with ice that changes on some records only."""
    cfg = synth.CONFIGS["t_py"]
    land, ice, lon, lat, z, sig = synth.static_fields(cfg)
    nrec = 5
    T, U, V, CI = [], [], [], []
    for step in range(1, nrec + 1):
        theta, p, u, v = synth.step_fields(cfg, step, land, z, lat)
        T.append(theta); U.append(u); V.append(v)
        CI.append(ice * (1.0 if step in (1, 2, 5) else 0.4))   # records 1,2 share ice; 3,4 share; 5 differs
    T, U, V, CI = (np.stack(x).astype(np.float32) for x in (T, U, V, CI))
    tt, sb, thc, ws, wd, fl = seabreezediag.diag_batched(1, land, z, sig, lon, lat, p, U, V, T, CI,
                                                          timestep=cfg.timestep, return_flags=True)
    assert tt == nrec + 1 and sb.shape == (nrec, cfg.ny, cfg.nx)
    ws_o = np.zeros_like(z); wd_o = np.zeros_like(z)
    for r in range(nrec):
        cd = oracle.py_get_dist(oracle.py_get_edges(land, CI[r]), land, lon, lat)
        o, f = oracle.py_diag(r + 1, p, z, sig, T[r], V[r], U[r], cd, ws_o, wd_o, timestep=cfg.timestep)
        assert_bits_equal(o[0], sb[r].astype(np.float32), "sb_con record %d" % r)
        assert np.array_equal(f, fl[r]), "flags record %d" % r
        ws_o, wd_o = o[2].copy(), o[3].copy()
    assert_bits_equal(ws_o, ws, "ws"); assert_bits_equal(wd_o, wd, "wd"); assert_bits_equal(o[1], thc, "t0")
    # no ice at all, and the per-record wrapper gives the same answer
    r1 = seabreezediag.diag_batched(1, land, z, sig, lon, lat, p, U, V, T, None, timestep=cfg.timestep)
    r2 = seabreezediag.diag(1, land, z, sig, lon, lat, p, U, V, T, None, timestep=cfg.timestep)
    assert_bits_equal(r1[1].astype(np.float32), r2[1].astype(np.float32), "batched vs per-record sb_con")
    assert_bits_equal(r1[3], r2[3], "batched vs per-record ws")


@pytest.mark.parametrize("name", ["t_gen", "c1"])
def test_generic_edges_dist(oracle, name):
    cfg = synth.CONFIGS[name]
    case = generic_case(cfg, oracle)
    h = cfg.h
    mask_g = np.zeros_like(case["coast_h"])
    m1 = np.zeros((cfg.ny, cfg.nx), np.uint8)
    generic.get_edges(mask_g, case["ice"], case["land"], h, swap=True, m1=m1)
    assert_bits_equal(case["coast_h"], mask_g, "get_edges (halo'd coast)")
    assert np.array_equal(m1, case["m1"]), "binarised mask m1"
    cd_g = np.empty_like(case["land"])
    generic.get_dist(mask_g, case["land"], case["lon"], case["lat"], 180.0, cd_g, h)
    cd_lit = oracle.get_dist(case["coast_h"], case["land"], case["lon"], case["lat"], 180.0, h, literal=True)
    assert_bits_equal(cd_lit, cd_g, "get_dist")
    f = synth.pad_halo(case["cdist"], h)
    generic.swap_bounds(f, h)
    assert_bits_equal(case["mask_h"], f, "swap_bounds")
    # literal no-op swap: caller-initialised halos are read, not written
    rng = np.random.default_rng(2)
    halo0 = (rng.random(mask_g.shape) > 0.5).astype(np.float32)
    mo, _ = oracle.get_edges(case["ice"], case["land"], h, swap=False, mask_in=halo0)
    mg = halo0.copy()
    generic.get_edges(mg, case["ice"], case["land"], h, swap=False)
    assert_bits_equal(mo, mg, "get_edges swap=False")


@pytest.mark.parametrize("name,level_scan,wk", [("t_gen", 0, 0), ("t_gen", 1, 0), ("c1", 0, 0), ("t_gen", 0, 2),
                                                 ("t_gen", 1, 2), ("c1", 0, 2), ("t_gen", 0, 3), ("t_gen", 1, 3),
                                                 ("c1", 0, 3), ("t_gen", 0, 1), ("c1", 0, 1)])
def test_generic_seabreeze_diag_steps(oracle, name, level_scan, wk):
    """wk = 0 / 1: the tile window kernel (k_trigger, the automatic choice on these small-radius grids);
    2 / 3: the strip window kernels (k_win4 / k_win8) forced."""
    with window_kernel(wk):
        _generic_seabreeze_diag_steps(oracle, name, level_scan)


def _generic_seabreeze_diag_steps(oracle, name, level_scan):
    cfg = synth.CONFIGS[name]
    case = generic_case(cfg, oracle)
    ny, nx = cfg.ny, cfg.nx
    so = [np.zeros((ny, nx), np.float32) for _ in range(4)]
    sg = [np.zeros((ny, nx), np.float32) for _ in range(4)]
    sp = [np.zeros((ny, nx), np.float32) for _ in range(4)]
    prm = generic._capi.default_params(SBD_VARIANT_GENERIC)
    prm.level_scan = level_scan
    plan = generic.Plan(SBD_VARIANT_GENERIC, nx, ny, cfg.nz, cfg.h, prm)
    plan.set_static(case["mask_h"], case["z_h"], case["sigma_h"])
    info = plan.info()
    assert info["nn_max"] >= 1 and info["n_coastal"] > 0
    trig = 0
    for step in (1, 2, 15, 16):
        theta_h, p, u, v = generic_step_inputs(cfg, case, step, oracle)
        fo = oracle.seabreeze_diag(cfg.timestep, step, p, u, v, theta_h, case["mask_h"], case["z_h"],
                                   case["sigma_h"], *so, cfg.h, literal=True, level_scan=level_scan)
        fg = np.zeros((ny, nx), np.uint8)
        generic.seabreeze_diag(cfg.timestep, step, p, u, v, theta_h, case["mask_h"], case["z_h"],
                               case["sigma_h"], *sg, params=prm, flags=fg)
        fp = np.zeros((ny, nx), np.uint8)
        plan.step(cfg.timestep, step, p, u, v, theta_h, *sp, flags=fp)
        for a, b, c, nm in zip(so, sg, sp, ("windspeed", "winddir", "thc", "sb_con")):
            assert_bits_equal(a, b, "one-shot %s step %d" % (nm, step))
            assert_bits_equal(a, c, "plan %s step %d" % (nm, step))
        assert np.array_equal(fo, fg) and np.array_equal(fo, fp), "flags step %d" % step
        trig += int((fo & 2).sum())
    assert trig > 0
    plan.close()


@pytest.mark.parametrize("nx,ny", [(97, 63), (130, 41), (33, 35)])
def test_generic_odd_sizes_scalar_scan(oracle, nx, ny):
    """nx*ny not a multiple of 4 (no aligned quads): the scalar level scan + k_final path, plan and
    one-shot, ragged tiles on both axes."""
    cfg = synth.Config("t_odd", nx, ny, 7, 4, "generic", 3.0, False, 1440.0, "odd-size test")
    case = generic_case(cfg, oracle)
    plan = generic.Plan(SBD_VARIANT_GENERIC, nx, ny, cfg.nz, cfg.h)
    plan.set_static(case["mask_h"], case["z_h"], case["sigma_h"])
    so = [np.zeros((ny, nx), np.float32) for _ in range(4)]
    sp = [np.zeros((ny, nx), np.float32) for _ in range(4)]
    sg = [np.zeros((ny, nx), np.float32) for _ in range(4)]
    for step in (1, 2, 15):
        theta_h, p, u, v = generic_step_inputs(cfg, case, step, oracle)
        fo = oracle.seabreeze_diag(cfg.timestep, step, p, u, v, theta_h, case["mask_h"], case["z_h"],
                                   case["sigma_h"], *so, cfg.h)
        fp = np.zeros((ny, nx), np.uint8)
        plan.step(cfg.timestep, step, p, u, v, theta_h, *sp, flags=fp)
        generic.seabreeze_diag(cfg.timestep, step, p, u, v, theta_h, case["mask_h"], case["z_h"], case["sigma_h"], *sg)
        for a, b, c, nm in zip(so, sp, sg, ("windspeed", "winddir", "thc", "sb_con")):
            assert_bits_equal(a, b, "plan %s step %d" % (nm, step))
            assert_bits_equal(a, c, "one-shot %s step %d" % (nm, step))
        assert np.array_equal(fo, fp)
    plan.close()


def test_generic_split_api_matches_single_call(oracle):
    """sbd_plan_step_scan + sbd_plan_step_finish (scan -> window -> final, the multi-GPU order) gives
    the same bits as sbd_plan_step (window -> scan + thresholds)."""
    import torch
    cfg = synth.CONFIGS["t_gen"]
    case = generic_case(cfg, oracle)
    ny, nx = cfg.ny, cfg.nx
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    plans = [generic.Plan(SBD_VARIANT_GENERIC, nx, ny, cfg.nz, cfg.h) for _ in range(2)]
    for pl in plans:
        pl.set_static(t(case["mask_h"]), t(case["z_h"]), t(case["sigma_h"]))
    st = [[torch.zeros((ny, nx), dtype=torch.float32, device=dev) for _ in range(4)] for _ in range(2)]
    so = [np.zeros((ny, nx), np.float32) for _ in range(4)]
    for step in (1, 2, 15):
        theta_h, p, u, v = generic_step_inputs(cfg, case, step, oracle)
        oracle.seabreeze_diag(cfg.timestep, step, p, u, v, theta_h, case["mask_h"], case["z_h"], case["sigma_h"], *so, cfg.h)
        dp, du, dv, dth = t(p), t(u), t(v), t(theta_h)
        plans[0].step(cfg.timestep, step, dp, du, dv, dth, *st[0])
        plans[1].step_scan(dp, du, dv)
        plans[1].step_finish(cfg.timestep, step, dth, *st[1])
        for pl in plans:
            pl.sync()
        for a, b, c, nm in zip(so, st[0], st[1], ("windspeed", "winddir", "thc", "sb_con")):
            assert_bits_equal(a, b.cpu().numpy(), "single %s step %d" % (nm, step))
            assert_bits_equal(a, c.cpu().numpy(), "split %s step %d" % (nm, step))


@pytest.mark.parametrize("name", ["t_gen", "c1"])
def test_stencil_byproducts(oracle, name):
    """the Sobel components px, py (G:350-357) out of get_edges, and out of the plan
    the radius nn at which the growing-window loop (G:188-214) stops with its counters n_l, n_s."""
    cfg = synth.CONFIGS[name]
    case = generic_case(cfg, oracle)
    ny, nx, h = cfg.ny, cfg.nx, cfg.h
    # the binarised, halo-swapped mask the Sobel reads: rebuild it on the oracle side from m1
    m1_h = synth.pad_halo(case["m1"].astype(np.float32), h)
    oracle.swap_bounds(m1_h, h)
    px_o, py_o = oracle.sobel(m1_h, h)
    mask = np.zeros((ny + 2 * h, nx + 2 * h), np.float32)
    px = np.full((ny, nx), 99.0, np.float32); py = np.full((ny, nx), 99.0, np.float32)
    generic.get_edges(mask, case["ice"], case["land"], h, px=px, py=py)
    assert_bits_equal(case["coast_h"], mask, "coast")
    assert_bits_equal(px_o, px, "px"); assert_bits_equal(py_o, py, "py")
    assert np.array_equal((px != 0) | (py != 0), mask[h:h + ny, h:h + nx] != 0)
    plan = generic.Plan(SBD_VARIANT_GENERIC, nx, ny, cfg.nz, h)
    plan.set_static(case["mask_h"], case["z_h"], case["sigma_h"])
    nn, nl, ns = plan.window_info()
    nn_o, nl_o, ns_o = oracle.window_info(case["mask_h"], h)
    assert np.array_equal(nn, nn_o) and np.array_equal(nl, nl_o) and np.array_equal(ns, ns_o)
    assert nn.max() == plan.info()["nn_max"] and int((nn > 0).sum()) == plan.info()["n_coastal"]
    plan.close()


def test_generic_virtual_halo(oracle):
    """Plan halo mode 7: swap_bounds folded into the kernels' reads -- theta arrives with garbage in
    its halo and the result equals the run with explicitly filled halos."""
    cfg = synth.CONFIGS["t_gen"]
    case = generic_case(cfg, oracle)
    ny, nx, h = cfg.ny, cfg.nx, cfg.h
    plan = generic.Plan(SBD_VARIANT_GENERIC, nx, ny, cfg.nz, h)
    plan.set_halo_mode(7)
    def junk(a):
        b = np.full_like(a, 1.0e30)
        b[h:h + ny, h:h + nx] = a[h:h + ny, h:h + nx]
        return b

    plan.set_static(junk(case["mask_h"]), junk(case["z_h"]), junk(case["sigma_h"]))
    so = [np.zeros((ny, nx), np.float32) for _ in range(4)]
    sp = [np.zeros((ny, nx), np.float32) for _ in range(4)]
    for step in (1, 2, 15):
        theta_h, p, u, v = generic_step_inputs(cfg, case, step, oracle)
        fo = oracle.seabreeze_diag(cfg.timestep, step, p, u, v, theta_h, case["mask_h"], case["z_h"],
                                   case["sigma_h"], *so, h)
        fp = np.zeros((ny, nx), np.uint8)
        plan.step(cfg.timestep, step, p, u, v, junk(theta_h), *sp, flags=fp)
        for a, b, nm in zip(so, sp, ("windspeed", "winddir", "thc", "sb_con")):
            assert_bits_equal(a, b, "%s step %d" % (nm, step))
        assert np.array_equal(fo, fp)
    plan.close()


def test_generic_pinned_host_buffers(oracle):
    """Host arrays in pinned memory: p,u,v are gathered in place over PCIe (zero-copy) for the southern
    rows while the DMA engine copies the northern rows (hybrid transfer) -- same bits as the oracle."""
    import torch
    cfg = synth.Config("t_pin", 256, 128, 10, 6, "generic", 5.0, False, 1440.0, "pinned-host test")
    case = generic_case(cfg, oracle)
    ny, nx = cfg.ny, cfg.nx
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    plan = generic.Plan(SBD_VARIANT_GENERIC, nx, ny, cfg.nz, cfg.h)
    plan.set_static(case["mask_h"], case["z_h"], case["sigma_h"])
    assert plan.info()["n_quads"] > 0
    so = [np.zeros((ny, nx), np.float32) for _ in range(4)]
    sp = [pin(np.zeros((ny, nx), np.float32)) for _ in range(4)]
    for step in (1, 2, 15):
        theta_h, p, u, v = generic_step_inputs(cfg, case, step, oracle)
        fo = oracle.seabreeze_diag(cfg.timestep, step, p, u, v, theta_h, case["mask_h"], case["z_h"],
                                   case["sigma_h"], *so, cfg.h)
        fp = torch.zeros((ny, nx), dtype=torch.uint8).pin_memory().numpy()
        plan.step(cfg.timestep, step, pin(p), pin(u), pin(v), pin(theta_h), *sp, flags=fp)
        for a, b, nm in zip(so, sp, ("windspeed", "winddir", "thc", "sb_con")):
            assert_bits_equal(a, b, "%s step %d" % (nm, step))
        assert np.array_equal(fo, fp)
    plan.close()


def test_generic_device_pointers_torch(oracle):
    """Same step with every array resident on the GPU (torch tensors): the path bench.py times."""
    import torch
    cfg = synth.CONFIGS["t_gen"]
    case = generic_case(cfg, oracle)
    ny, nx = cfg.ny, cfg.nx
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    plan = generic.Plan(SBD_VARIANT_GENERIC, nx, ny, cfg.nz, cfg.h)
    plan.set_static(t(case["mask_h"]), t(case["z_h"]), t(case["sigma_h"]))
    so = [np.zeros((ny, nx), np.float32) for _ in range(4)]
    sg = [torch.zeros((ny, nx), dtype=torch.float32, device=dev) for _ in range(4)]
    fg = torch.zeros((ny, nx), dtype=torch.uint8, device=dev)
    for step in (1, 2):
        theta_h, p, u, v = generic_step_inputs(cfg, case, step, oracle)
        fo = oracle.seabreeze_diag(cfg.timestep, step, p, u, v, theta_h, case["mask_h"], case["z_h"],
                                   case["sigma_h"], *so, cfg.h)
        plan.step(cfg.timestep, step, t(p), t(u), t(v), t(theta_h), *sg, flags=fg)
        plan.sync()
        for a, b, nm in zip(so, sg, ("windspeed", "winddir", "thc", "sb_con")):
            assert_bits_equal(a, b.cpu().numpy(), "%s step %d" % (nm, step))
        assert np.array_equal(fo, fg.cpu().numpy())
    plan.close()


@pytest.mark.parametrize("wk", [0, 1, 2, 3, 38, 316, 201, 301])
@pytest.mark.parametrize("nx,ny,h,filt,expect", [(160, 96, 24, 14.0, "big tile"), (320, 240, 54, 40.0, "big tile"),
                                                   (300, 200, 72, 60.0, "global path")])
def test_wide_windows_km_scale(oracle, nx, ny, h, filt, expect, wk):
    """km-scale regional grids (config 3): the final search radius reaches tens of cells.  wk = 0: the
    8-cell strip kernel (k_win8, radii up to 56 in shared memory, beyond that the global-memory path);
    wk = 1: the tile kernel (k_trigger, 128x64 tile with patch-raster order); wk = 2: round 1's 4-cell
    strip kernel (k_win4)."""
    cfg = synth.Config("t_big", nx, ny, 4, h, "generic", filt, True, 1440.0, "wide-window test")
    case = generic_case(cfg, oracle)
    plan = generic.Plan(SBD_VARIANT_GENERIC, nx, ny, cfg.nz, cfg.h)
    if wk == 38:     # k_win8 with 8-cell strips
        plan.set_window_kernel(3)
        plan.set_strip_cells(8)
    elif wk == 316:  # k_win8 with 64 x 16 tiles
        plan.set_window_kernel(3)
        plan.set_tile_rows(16)
    elif wk in (201, 301):   # k_win4 / k_win8 staging from the per-step x-plane (k_xplane) instead of theta, gzsm, code
        plan.set_window_kernel(wk // 100)
        plan.set_xplane(True)
    else:
        plan.set_window_kernel(wk)
    plan.set_static(case["mask_h"], case["z_h"], case["sigma_h"])
    nn_max = plan.info()["nn_max"]
    assert nn_max > 16, plan.info()
    if expect == "global path":
        assert nn_max > 56, plan.info()
    so = [np.zeros((ny, nx), np.float32) for _ in range(4)]
    sp = [np.zeros((ny, nx), np.float32) for _ in range(4)]
    for step in (1, 2):
        theta_h, p, u, v = generic_step_inputs(cfg, case, step, oracle)
        fo = oracle.seabreeze_diag(cfg.timestep, step, p, u, v, theta_h, case["mask_h"], case["z_h"], case["sigma_h"],
                                   *so, cfg.h, literal=False)
        fp = np.zeros((ny, nx), np.uint8)
        plan.step(cfg.timestep, step, p, u, v, theta_h, *sp, flags=fp)
        for a, b, nm in zip(so, sp, ("windspeed", "winddir", "thc", "sb_con")):
            assert_bits_equal(a, b, "%s step %d" % (nm, step))
        assert np.array_equal(fo, fp)
    plan.close()


@pytest.mark.parametrize("wk", [3, 2, 301, 201])
@pytest.mark.parametrize("what", ["negative", "nan", "zero"])
def test_wide_windows_signed_class_fallback(oracle, what, wk):
    """k_win8 keeps x = +-t0 in shared memory, which needs t0 >= +0; a tile that stages a negative or
    NaN t0 must fall back to the per-cell global path and still give the oracle's bits (NaNs included).
    t0 == 0 (both signs) stays on the fast path."""
    cfg = synth.Config("t_neg", 200, 130, 3, 30, "generic", 18.0, True, 1440.0, "signed-class fallback test")
    case = generic_case(cfg, oracle)
    ny, nx, h = cfg.ny, cfg.nx, cfg.h
    plan = generic.Plan(SBD_VARIANT_GENERIC, nx, ny, cfg.nz, h)
    plan.set_window_kernel(wk // 100 if wk > 100 else wk)    # k_win8, k_win4; 301 / 201: staged from the x-plane, whose
    plan.set_xplane(wk > 100)                                 # NaN marker sends the tile to the global path
    if what == "nan":
        plan.set_tile_rows(16)
    plan.set_static(case["mask_h"], case["z_h"], case["sigma_h"])
    assert plan.info()["nn_max"] >= 12, plan.info()
    so = [np.zeros((ny, nx), np.float32) for _ in range(4)]
    sp = [np.zeros((ny, nx), np.float32) for _ in range(4)]
    for step in (1, 2):
        theta_h, p, u, v = generic_step_inputs(cfg, case, step, oracle)
        if what == "negative":
            theta_h[h + 40:h + 44, h + 70:h + 90] = -5.0      # one patch of cells below zero
            theta_h[h + 100, h + 150] = -300.0
        elif what == "nan":
            theta_h[h + 64, h + 100] = np.nan
        else:
            theta_h[h + 30:h + 60, h + 20:h + 120] = case["z_h"][h + 30:h + 60, h + 20:h + 120] * 0.0   # t0 = +0 over sea
            theta_h[h + 61, h + 30] = -0.0
        fo = oracle.seabreeze_diag(cfg.timestep, step, p, u, v, theta_h, case["mask_h"], case["z_h"], case["sigma_h"],
                                   *so, cfg.h, literal=False)
        fp = np.zeros((ny, nx), np.uint8)
        plan.step(cfg.timestep, step, p, u, v, theta_h, *sp, flags=fp)
        for a, b, nm in zip(so, sp, ("windspeed", "winddir", "thc", "sb_con")):
            assert_bits_equal(a, b, "%s %s step %d" % (what, nm, step))
        assert np.array_equal(fo, fp)
    plan.close()


def test_determinism_run_twice(oracle):
    cfg = synth.CONFIGS["c5"]
    land, ice, lon, lat, z, sig = synth.static_fields(cfg)
    a = seabreeze.get_dist(seabreeze.get_edges(F(land), F(ice)), F(land), lon, lat)
    b = seabreeze.get_dist(seabreeze.get_edges(F(land), F(ice)), F(land), lon, lat)
    assert_bits_equal(a, b, "get_dist determinism")
