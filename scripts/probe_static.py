"""Timing probe of the static chain (get_edges -> get_dist -> set_static) on a named config, device-resident,
CUDA events on the default stream; generic configs use the generic entry points, python configs sbd_py_*."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seabreeze_param_b200 import generic, synth, _capi
import ctypes as C

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
cfg = synth.CONFIGS[name]
dev = torch.device("cuda:0")
h, nx, ny, nz = cfg.h, cfg.nx, cfg.ny, cfg.nz
land, ice, lon, lat, z, sig = synth.static_fields(cfg)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
dl, di, dlon, dlat, dz, ds = map(t, (land, ice, lon, lat, z, sig))
lib = _capi.load()

def ev_time(fn, reps=3):
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return out

if cfg.variant == "generic":
    mask_h = torch.zeros((ny + 2 * h, nx + 2 * h), dtype=torch.float32, device=dev)
    cdist = torch.empty((ny, nx), dtype=torch.float32, device=dev)
    print(name, "get_edges ms", ev_time(lambda: generic.get_edges(mask_h, di, dl, h)))
    print(name, "get_dist ms (k=%d)" % h, ev_time(lambda: generic.get_dist(mask_h, dl, dlon, dlat, 180.0, cdist, h)))
    def haloed(a2d):
        f = torch.zeros((ny + 2 * h, nx + 2 * h), dtype=torch.float32, device=dev); f[h:h + ny, h:h + nx] = a2d
        generic.swap_bounds(f, h); return f
    mk, zh, sh = haloed(cdist), haloed(dz), haloed(ds)
    plan = generic.Plan(generic.SBD_VARIANT_GENERIC, nx, ny, nz, h)
    def ss():
        t0 = time.perf_counter(); plan.set_static(mk, zh, sh); torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t0)
    print(name, "set_static ms (wall)", [ss() for _ in range(3)], plan.info())
else:
    coast = torch.empty((ny, nx), dtype=torch.float32, device=dev)
    cdist = torch.empty((ny, nx), dtype=torch.float32, device=dev)
    k = C.c_int(0)
    q = _capi.ptr
    print(name, "py get_edges ms", ev_time(lambda: _capi.check(lib.sbd_py_get_edges_f32(q(dl), q(di), nx, ny, q(coast)))))
    print(name, "py get_dist ms", ev_time(lambda: _capi.check(lib.sbd_py_get_dist_f32(q(coast), q(dl), q(dlon), q(dlat), nx, ny, 180.0, q(cdist), C.byref(k)))), "k", k.value)
    plan = generic.Plan(generic.SBD_VARIANT_PYTHON, nx, ny, nz)
    def ss():
        t0 = time.perf_counter(); plan.set_static(cdist, dz, ds); torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t0)
    print(name, "set_static ms (wall)", [ss() for _ in range(3)], plan.info())
