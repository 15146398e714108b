"""Timing probe of the generic plan step on a named config with device-generated 3-D fields (2-D fields from
synth, so masks / radii are the benchmark's): per-kernel times of every window-kernel form.
usage: probe_win.py <cfg> [modes e.g. 3,2] [nz override]"""
import sys, os, time, math
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seabreeze_param_b200 import generic, synth
from seabreeze_param_b200._capi import SBD_VARIANT_GENERIC, SBD_VARIANT_PYTHON

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
modes = [int(m) for m in (sys.argv[2] if len(sys.argv) > 2 else "3,2").split(",")]
cfg = synth.CONFIGS[name]
nz = int(sys.argv[3]) if len(sys.argv) > 3 else cfg.nz
dev = torch.device("cuda:0")
h, nx, ny = cfg.h, cfg.nx, cfg.ny
land, ice, lon, lat, z, sig = synth.static_fields(cfg)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
g = torch.Generator(device=dev); g.manual_seed(1)
latr = torch.deg2rad(t(lat).double())[:, None]
theta = (300.0 - 30.0 * torch.sin(latr) ** 2 + 3.0 * t(land).double() + 0.5 * torch.randn((ny, nx), generator=g, device=dev, dtype=torch.float64)).float()
u = 6.0 * torch.randn((nz, ny, nx), generator=g, device=dev)
v = 6.0 * torch.randn((nz, ny, nx), generator=g, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
if cfg.variant == "generic":
    k = torch.arange(nz, device=dev, dtype=torch.float32)[:, None, None]
    p = (101325.0 * torch.exp(-t(z) / 8000.0)[None] * torch.exp(-k * math.log(100.0) / max(nz - 1, 1))).contiguous()
    mask_h = torch.zeros((ny + 2 * h, nx + 2 * h), dtype=torch.float32, device=dev)
    generic.get_edges(mask_h, t(ice), t(land), h)
    cdist = torch.empty((ny, nx), dtype=torch.float32, device=dev)
    generic.get_dist(mask_h, t(land), t(lon), t(lat), 180.0, cdist, h)
    def haloed(a2d):
        f = torch.zeros((ny + 2 * h, nx + 2 * h), dtype=torch.float32, device=dev); f[h:h + ny, h:h + nx] = a2d
        generic.swap_bounds(f, h); return f
    mk, zh, sh, th = haloed(cdist), haloed(t(z)), haloed(t(sig)), haloed(theta)
    for wk in modes:
        plan = generic.Plan(SBD_VARIANT_GENERIC, nx, ny, nz, h)
        plan.set_window_kernel(wk)
        if os.environ.get("PROBE_SPLIT"):
            plan.set_tile_split(int(os.environ["PROBE_SPLIT"]))
        if os.environ.get("PROBE_ROWS"):
            plan.set_tile_rows(int(os.environ["PROBE_ROWS"]))
        plan.set_static(mk, zh, sh)
        if os.environ.get("PROBE_TILES"):
            # how far from a perfectly divisible workload is this grid?  window elements per 64 x 32 tile, greedy
            # heaviest-first assignment to 296 CTA slots (148 SMs x 2) against the mean load per slot
            import heapq
            nn = plan.window_info()[0].astype(np.float64)
            el = (2 * nn + 1) ** 2 * (nn > 0)
            ty, tx = (ny + 31) // 32, (nx + 63) // 64
            pad = np.zeros((ty * 32, tx * 64)); pad[:ny, :nx] = el
            tc = pad.reshape(ty, 32, tx, 64).sum(axis=(1, 3)).ravel()
            tc = np.sort(tc[tc > 0])[::-1]
            slots = [0.0] * 296
            heapq.heapify(slots)
            for c in tc:
                heapq.heappush(slots, heapq.heappop(slots) + c)
            print("   tiles with work %d, window elements %.4g, heaviest tile %.4g, mean per slot %.4g, greedy makespan %.4g (x%.2f of mean)"
                  % (len(tc), tc.sum(), tc[0], tc.sum() / 296, max(slots), max(slots) / (tc.sum() / 296)))
        st = [torch.zeros((ny, nx), dtype=torch.float32, device=dev) for _ in range(4)]
        for s in range(3):
            plan.step(cfg.timestep, s + 1, p, u, v, th, *st)
        plan.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for s in range(6):
            flush.fill_(float(s)); torch.cuda.synchronize()
            t0 = time.perf_counter(); plan.step(cfg.timestep, s + 4, p, u, v, th, *st); plan.sync(); ts.append(1e6 * (time.perf_counter() - t0))
        plan.set_timing(True)
        ks = []
        for s in range(4):
            flush.fill_(float(s)); plan.step(cfg.timestep, s + 10, p, u, v, th, *st); ks.append(plan.last_timing())
        print(name, "window kernel mode", wk, plan.info()["nn_max"], "step wall us:", " ".join("%.0f" % x for x in ts),
              "| per-kernel (window, scan, final) us:", " ".join("(%.0f,%.0f,%.0f)" % (a * 1e3, b * 1e3, c * 1e3) for a, b, c in ks))
        print("   checksum", float(st[3].double().sum()), float(st[2].double().abs().sum()))
        plan.close()
else:
    import ctypes as C
    from seabreeze_param_b200 import _capi
    lib = _capi.load(); q = _capi.ptr
    p = t(synth.pressure_1d(nz))
    coast = torch.empty((ny, nx), dtype=torch.float32, device=dev); cdist = torch.empty_like(coast); kk = C.c_int(0)
    dl, di, dlon, dlat = t(land), t(ice), t(lon), t(lat)     # named: a temporary would be freed before the call reads it
    _capi.check(lib.sbd_py_get_edges_f32(q(dl), q(di), nx, ny, q(coast)))
    _capi.check(lib.sbd_py_get_dist_f32(q(coast), q(dl), q(dlon), q(dlat), nx, ny, 180.0, q(cdist), C.byref(kk)))
    like = os.environ.get("PROBE_BENCHLIKE", "")     # "w": coherent winds (bench.gen_uv), "f": flag plane, "s": state threaded
    kw = {}
    if "w" in like:
        import bench
        ctx = type("Ctx", (), {"torch": torch, "dev": dev})()
        u = bench.gen_uv(ctx, nz, ny, nx, 0, ny, 100, 0); v = bench.gen_uv(ctx, nz, ny, nx, 0, ny, 100, 1)
    if "f" in like:
        kw["flags"] = torch.zeros((ny, nx), dtype=torch.uint8, device=dev)
    recs = None
    if "r" in like:                                  # a different record (theta, u, v) every step, as bench_c4 does
        import bench
        ctx = type("Ctx", (), {"torch": torch, "dev": dev})()
        nrec = int(os.environ.get("PROBE_NREC", "6"))
        recs = [(t(bench.theta_host(cfg, 100 + r, land, lat, hour=float(r))), bench.gen_uv(ctx, nz, ny, nx, 0, ny, 100 + r, 0),
                 bench.gen_uv(ctx, nz, ny, nx, 0, ny, 100 + r, 1)) for r in range(nrec)]
    for wk in modes:
        plan = generic.Plan(SBD_VARIANT_PYTHON, nx, ny, nz)
        plan.set_window_kernel(wk)
        if os.environ.get("PROBE_STRIP8") and wk == 3:
            plan.set_strip_cells(8)
        if "p" in like:                              # the plan as bench_c4 builds it: row window, statistics handed in, own stream
            st4 = np.zeros(4, np.float32)
            generic.sigmoid(t(sig), stats=st4)
            plan.set_row_window(0, ny)
            stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); plan.set_stream(stream.cuda_stream)
            plan.set_static(cdist, t(z), t(sig), stats=st4[:3].copy())
        else:
            plan.set_static(cdist, t(z), t(sig))
        print("   plan info", {k: plan.info()[k] for k in ("nn_max", "n_coastal")})
        ws = torch.zeros((ny, nx), dtype=torch.float32, device=dev); wd = torch.zeros_like(ws)
        out = torch.zeros((4, ny, nx), dtype=torch.float32, device=dev)
        out2 = torch.zeros((4, ny, nx), dtype=torch.float32, device=dev)
        nstep = [0]
        def one(stepno):
            global theta, u, v
            if recs is not None:
                theta, u, v = recs[nstep[0] % len(recs)]
            if "s" in like:
                o, pv = (out, out2) if nstep[0] % 2 == 0 else (out2, out)
                plan.step_py(stepno, cfg.timestep, p, u, v, theta, pv[2], pv[3], o, **kw)
            else:
                plan.step_py(stepno, cfg.timestep, p, u, v, theta, ws, wd, out, **kw)
            nstep[0] += 1
        for s in range(3):
            one(s + 1)
        plan.sync()
        ts = []
        for s in range(6):
            flush.fill_(float(s)); torch.cuda.synchronize()
            t0 = time.perf_counter(); one(s + 4); plan.sync(); ts.append(1e6 * (time.perf_counter() - t0))
        plan.set_timing(True)
        ks = []
        for s in range(4):
            flush.fill_(float(s)); one(s + 10); ks.append(plan.last_timing()[0])
        if "f" in like:
            print("   trigger cells", int(((kw["flags"] & 2) > 0).sum().item()))
        print(name, "window kernel mode", wk, plan.info()["nn_max"], "step wall us:", " ".join("%.0f" % x for x in ts),
              "| kernel us:", " ".join("%.0f" % (a * 1e3) for a in ks))
        print("   checksum", float(out[0][out[0] < 1e19].double().sum()), float(out[1].double().sum()))
        plan.close()
