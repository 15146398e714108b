"""Ground-truth timing probe of the plan step: wall clock over many steps, per-step events with an
L2 flush in between, and different numbers of rotating input sets."""
import sys, time, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seabreeze_param_b200 import generic, synth, halo
from seabreeze_param_b200._capi import SBD_VARIANT_GENERIC, load
import bench

cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
dev = torch.device("cuda:0")
h, nx, ny, nz = cfg.h, cfg.nx, cfg.ny, cfg.nz
land, ice, lon, lat, z, sig, steps = bench.build_case(cfg, 0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
mask_h = torch.zeros((ny + 2 * h, nx + 2 * h), dtype=torch.float32, device=dev)
generic.get_edges(mask_h, t(ice), t(land), h)
cdist = torch.empty((ny, nx), dtype=torch.float32, device=dev)
generic.get_dist(mask_h, t(land), t(lon), t(lat), 180.0, cdist, h)
def haloed(a2d):
    f = torch.zeros((ny + 2 * h, nx + 2 * h), dtype=torch.float32, device=dev); f[h:h + ny, h:h + nx] = a2d; return f
ex = halo.RowBlockHalo(nx, ny, h, 0, 1, dev)
mk, zh, sh = ex.swap_bounds(haloed(cdist)), ex.swap_bounds(haloed(t(z))), ex.swap_bounds(haloed(t(sig)))
dsteps = [(ex.swap_bounds(haloed(t(a))), t(b), t(c), t(d)) for a, b, c, d in steps]
state = [torch.zeros((ny, nx), dtype=torch.float32, device=dev) for _ in range(4)]
plan = generic.Plan(SBD_VARIANT_GENERIC, nx, ny, nz, h)
stream = torch.cuda.Stream(device=dev)
torch.cuda.synchronize()
torch.cuda.set_stream(stream)
plan.set_stream(stream.cuda_stream)
plan.set_static(mk, zh, sh)
print(plan.info())
flush = torch.empty(512 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
flags = torch.zeros((ny, nx), dtype=torch.uint8, device=dev) if os.environ.get("PROBE_FLAGS") else None
def step(k):
    th, p, u, v = dsteps[k % len(dsteps)]
    plan.step(cfg.timestep, k + 2, p, u, v, th, *state, flags=flags)
for k in range(5): step(k)
torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(200): step(k)
torch.cuda.synchronize()
print("wall: %.2f us/step (back to back, %d sets)" % ((time.perf_counter() - t0) / 200 * 1e6, len(dsteps)))
t0 = time.perf_counter()
for k in range(200): step(k)
t1 = time.perf_counter()
torch.cuda.synchronize()
print("host issue: %.2f us/step" % ((t1 - t0) / 200 * 1e6))
ts = []
for k in range(20):
    flush.fill_(float(k))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); step(k); e1.record(stream)
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print("flushed per-step events (us):", " ".join("%.1f" % x for x in ts))
ts = []
for k in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); step(0); e1.record(stream)
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print("same-set per-step events (us):", " ".join("%.1f" % x for x in ts))
plan.set_timing(True)
ts = []
for k in range(12):
    flush.fill_(float(k)); step(k); ts.append(plan.last_timing())
print("per-kernel (window, scan, join+final) us:", " ".join("(%.1f,%.1f,%.1f)" % (a * 1e3, b * 1e3, c * 1e3) for a, b, c in ts))
