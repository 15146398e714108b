"""Per-CTA timeline of the wide-window kernel (k_win8) on a generic config: which SM ran which tile when.
Needs the diagnostic build  seabreeze_param_b200/_lib/libsbd_b200_trace.so  (python scripts/probe_tiles.py --build,
run where nvcc is; the .so travels to the GPU box).  usage: probe_tiles.py <cfg> [split]"""
import sys, os, subprocess, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from seabreeze_param_b200 import build as B
TRACE = os.path.join(B.OUT_DIR, "libsbd_b200_trace.so")
if "--build" in sys.argv:
    subprocess.check_call(["nvcc"] + B.NVCC_FLAGS + ["-DSBD_TRACE_TILES", B.SRC, "-o", TRACE])
    sys.exit(0)
B.OUT = TRACE                       # load the diagnostic library instead of the product one
B.needs_build = lambda: False
import numpy as np, torch
from seabreeze_param_b200 import generic, synth, _capi
from seabreeze_param_b200._capi import SBD_VARIANT_GENERIC
name = sys.argv[1] if len(sys.argv) > 1 else "c3r8"
split = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = synth.CONFIGS[name]
dev = torch.device("cuda:0")
h, nx, ny, nz = cfg.h, cfg.nx, cfg.ny, 4
land, ice, lon, lat, z, sig = synth.static_fields(cfg)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
g = torch.Generator(device=dev); g.manual_seed(1)
theta = (300.0 + 3.0 * t(land) + 0.5 * torch.randn((ny, nx), generator=g, device=dev)).float()
u = 6.0 * torch.randn((nz, ny, nx), generator=g, device=dev); v = 6.0 * torch.randn((nz, ny, nx), generator=g, device=dev)
k = torch.arange(nz, device=dev, dtype=torch.float32)[:, None, None]
p = (101325.0 * torch.exp(-t(z) / 8000.0)[None] * torch.exp(-k * 1.5)).contiguous()
mask_h = torch.zeros((ny + 2 * h, nx + 2 * h), dtype=torch.float32, device=dev)
generic.get_edges(mask_h, t(ice), t(land), h)
cdist = torch.empty((ny, nx), dtype=torch.float32, device=dev)
generic.get_dist(mask_h, t(land), t(lon), t(lat), 180.0, cdist, h)
def haloed(a2d):
    f = torch.zeros((ny + 2 * h, nx + 2 * h), dtype=torch.float32, device=dev); f[h:h + ny, h:h + nx] = a2d
    generic.swap_bounds(f, h); return f
mk, zh, sh, th = haloed(cdist), haloed(t(z)), haloed(t(sig)), haloed(theta)
plan = generic.Plan(SBD_VARIANT_GENERIC, nx, ny, nz, h)
plan.set_window_kernel(3); plan.set_tile_split(split)
if os.environ.get("PROBE_ROWS"):
    plan.set_tile_rows(int(os.environ["PROBE_ROWS"]))
plan.set_static(mk, zh, sh)
nn = plan.window_info()[0].astype(np.float64)
el = (2 * nn + 1) ** 2 * (nn > 0)
TR = int(os.environ.get("PROBE_ROWS", "32"))
plan.set_tile_rows(TR)              # (again: explicit even without the variable, so the cost map matches the launch)
ty, tx = (ny + TR - 1) // TR, (nx + 63) // 64
pad = np.zeros((ty * TR, tx * 64)); pad[:ny, :nx] = el
tcost = pad.reshape(ty, TR, tx, 64).sum(axis=(1, 3)).ravel()
st = [torch.zeros((ny, nx), dtype=torch.float32, device=dev) for _ in range(4)]
for s in range(4):
    plan.step(cfg.timestep, s + 1, p, u, v, th, *st)
plan.sync()
lib = _capi.load()
n = 16384
buf = np.zeros((n, 4), np.uint64)
rc = lib.sbd_debug_tile_trace(C.c_void_p(buf.ctypes.data), n)
assert rc == 0
buf = buf[buf[:, 3] > 0]
ent, sm, t0, t1 = buf[:, 0].astype(np.int64), buf[:, 1].astype(np.int64), buf[:, 2].astype(np.float64), buf[:, 3].astype(np.float64)
tile, part, nparts = ent & 0xfffff, (ent >> 20) & 0x3f, ent >> 26
T0 = t0.min(); t0 = (t0 - T0) / 1e3; t1 = (t1 - T0) / 1e3
dur = t1 - t0
cost = tcost[tile] / np.maximum(nparts, 1)
print("rows per tile %d;" % TR, "%s split %d: %d CTAs, kernel span %.0f us, sum of CTA time %.0f us = %.1f slots busy on average (of 296)"
      % (name, split, len(buf), t1.max(), dur.sum(), dur.sum() / t1.max()))
o = np.argsort(-dur)
print("longest CTAs: dur us / start us / window elements / elements per us / sm")
for i in o[:12]:
    print("  %7.0f %7.0f %10.3g %8.0f  sm %3d  tile %5d part %d/%d" % (dur[i], t0[i], cost[i], cost[i] / dur[i], sm[i], tile[i], part[i], nparts[i]))
big = cost > 0.5 * cost.max()
print("CTAs above half the heaviest cost: %d, mean rate %.0f el/us; all CTAs with work: mean rate %.0f el/us"
      % (big.sum(), (cost[big] / dur[big]).mean(), (cost[cost > 0].sum() / dur[cost > 0].sum())))
edges = np.linspace(0, t1.max(), 11)
act = [(np.minimum(t1, b) - np.maximum(t0, a)).clip(0).sum() / (b - a) for a, b in zip(edges[:-1], edges[1:])]
print("active CTAs per tenth of the span:", " ".join("%.0f" % x for x in act))
light = cost < 1e4
print("CTAs with (almost) no window work: %d, mean duration %.1f us" % (light.sum(), dur[light].mean() if light.any() else 0))
