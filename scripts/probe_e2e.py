"""Probe: where does the host-buffer (e2e) step spend its time?"""
import sys, time, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seabreeze_param_b200 import generic, synth
from seabreeze_param_b200._capi import SBD_VARIANT_GENERIC
import bench
cfg = synth.CONFIGS["c2"]
h, nx, ny, nz = cfg.h, cfg.nx, cfg.ny, cfg.nz
land, ice, lon, lat, z, sig, steps = bench.build_case(cfg, 0)
mask_h = np.zeros((ny + 2 * h, nx + 2 * h), np.float32)
generic.get_edges(mask_h, ice, land, h)
cdist = np.empty((ny, nx), np.float32)
generic.get_dist(mask_h, land, lon, lat, 180.0, cdist, h)
mk = generic.swap_bounds(synth.pad_halo(cdist, h), h)
zh = generic.swap_bounds(synth.pad_halo(z, h), h)
sh = generic.swap_bounds(synth.pad_halo(sig, h), h)
plan = generic.Plan(SBD_VARIANT_GENERIC, nx, ny, nz, h)
plan.set_static(mk, zh, sh)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
theta, p, u, v = steps[0]
th_h = generic.swap_bounds(synth.pad_halo(theta, h), h)
for label, conv in (("pageable", lambda a: np.ascontiguousarray(a)), ("pinned", pin)):
    P, U, V, T = conv(p), conv(u), conv(v), conv(th_h)
    st = [conv(np.zeros((ny, nx), np.float32)) for _ in range(4)]
    for i in range(2):
        plan.step(cfg.timestep, i + 2, P, U, V, T, *st)
    ts = []
    for i in range(5):
        t0 = time.perf_counter()
        plan.step(cfg.timestep, i + 4, P, U, V, T, *st)
        ts.append((time.perf_counter() - t0) * 1e3)
    print(label, "plan.step host ms:", " ".join("%.2f" % x for x in ts))
d = torch.empty(p.size, dtype=torch.float32, device="cuda")
hp = torch.from_numpy(p.ravel()).pin_memory()
torch.cuda.synchronize()
t0 = time.perf_counter(); d.copy_(hp, non_blocking=True); torch.cuda.synchronize()
print("torch pinned H2D of p (%.0f MB): %.2f ms" % (p.nbytes / 1e6, (time.perf_counter() - t0) * 1e3))
