"""Timing probe of the python-variant step (diag, seabreeze_diag_python.f90:49-285) on a device-resident
plan: sbd_plan_step_py back to back over rotating input sets, tile window kernel vs strip window kernel.
usage: probe_py.py [config=c4] [nsets=3]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seabreeze_param_b200 import generic, seabreeze, synth
from seabreeze_param_b200._capi import SBD_VARIANT_PYTHON

cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c4"]
nsets = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
nx, ny, nz = cfg.nx, cfg.ny, cfg.nz
F = lambda a: np.asfortranarray(a.T) if a.ndim == 2 else np.asfortranarray(a.transpose(2, 1, 0))
land, ice, lon, lat, z, sig = synth.static_fields(cfg)
t0 = time.perf_counter()
coast = seabreeze.get_edges(F(land), F(ice))
cd = seabreeze.get_dist(coast, F(land), lon, lat)
print("get_edges + get_dist (host arrays): %.1f ms" % (1e3 * (time.perf_counter() - t0)))
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
cd_c = np.ascontiguousarray(np.asarray(cd).T)   # (ny, nx) C-order == Fortran (nx, ny)
sets = []
for s in range(nsets):
    theta, p, u, v = synth.step_fields(cfg, s + 1, land, z, lat)
    sets.append((t(theta), t(p), t(u), t(v)))
ws = torch.zeros((ny, nx), dtype=torch.float32, device=dev)
wd = torch.zeros((ny, nx), dtype=torch.float32, device=dev)
out = torch.zeros((4, ny, nx), dtype=torch.float32, device=dev)
for wk in (1, 0):
    plan = generic.Plan(SBD_VARIANT_PYTHON, nx, ny, nz)
    plan.set_window_kernel(wk)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    plan.set_stream(stream.cuda_stream)
    t0 = time.perf_counter()
    plan.set_static(t(cd_c), t(z), t(sig))
    plan.sync()
    print("wk=%d set_static %.1f ms" % (wk, 1e3 * (time.perf_counter() - t0)), plan.info())
    def step(k):
        th, p, u, v = sets[k % nsets]
        plan.step_py(k + 2, cfg.timestep, p, u, v, th, ws, wd, out)
    for k in range(3): step(k)
    plan.sync()
    n = 24
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(n): step(k)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("wk=%d (%s window kernel): %.1f us/step, %.3e cells/s, %.0f GB/s of 48 B/cell" % (
        wk, "tile" if wk == 1 else "auto/strip", 1e3 * ms, nx * ny / (ms * 1e-3), 48.0 * nx * ny / (ms * 1e-3) / 1e9))
    chk = float(out[0].nan_to_num(0.0, 0.0, 0.0).clamp(-1e3, 1e3).double().sum().item())
    print("   checksum sb_con %.9e" % chk)
    plan.close()
