"""torchrun --nproc-per-node N scripts/check_multigpu.py : N-rank row-block decomposition (NCCL halo
exchange, chained global statistics) must be bit-identical to the single-domain run."""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seabreeze_param_b200 import generic, halo, synth
from seabreeze_param_b200._capi import SBD_VARIANT_GENERIC

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
nyl = 48
cfg = synth.Config("mg", 256, nyl * world, 10, 6, "generic", 5.0, False, 1440.0, "multi-gpu check")
h, nx, ny, nz = cfg.h, cfg.nx, cfg.ny, cfg.nz
land, ice, lon, lat, z, sig = synth.static_fields(cfg, seed=3)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
def haloed(a2d, n_rows):
    f = torch.zeros((n_rows + 2 * h, nx + 2 * h), dtype=torch.float32, device=dev); f[h:h + n_rows, h:h + nx] = a2d; return f
# static chain on the global grid (every rank; get_dist is static and cheap here)
mask_h = torch.zeros((ny + 2 * h, nx + 2 * h), dtype=torch.float32, device=dev)
generic.get_edges(mask_h, t(ice), t(land), h)
cdist = torch.empty((ny, nx), dtype=torch.float32, device=dev)
generic.get_dist(mask_h, t(land), t(lon), t(lat), 180.0, cdist, h)
r0, r1 = rank * nyl, (rank + 1) * nyl
ex = halo.RowBlockHalo(nx, nyl, h, rank, world, dev)
mk = ex.swap_bounds(haloed(cdist[r0:r1], nyl)); zh = ex.swap_bounds(haloed(t(z[r0:r1]), nyl)); sh = ex.swap_bounds(haloed(t(sig[r0:r1]), nyl))
stats = ex.global_sigma_stats(sh)
plan = generic.Plan(SBD_VARIANT_GENERIC, nx, nyl, nz, h)
plan.set_static(mk, zh, sh, stats=stats)
st = [torch.zeros((nyl, nx), dtype=torch.float32, device=dev) for _ in range(4)]
fl = torch.zeros((nyl, nx), dtype=torch.uint8, device=dev)
# second plan: no exchange at all -- halo rows of theta are read from the neighbours' planes (peer memory)
peer = halo.PeerHalo(1, nyl, nx, h, rank, world, dev)
plan_p = generic.Plan(SBD_VARIANT_GENERIC, nx, nyl, nz, h)
plan_p.set_halo_mode(ex.halo_mode())
plan_p.set_static(mk, zh, sh, stats=stats)
st_p = [torch.zeros((nyl, nx), dtype=torch.float32, device=dev) for _ in range(4)]
ok = True
if rank == 0:
    ex1 = halo.RowBlockHalo(nx, ny, h, 0, 1, dev)
    mk1 = ex1.swap_bounds(haloed(cdist, ny)); zh1 = ex1.swap_bounds(haloed(t(z), ny)); sh1 = ex1.swap_bounds(haloed(t(sig), ny))
    plan1 = generic.Plan(SBD_VARIANT_GENERIC, nx, ny, nz, h)
    plan1.set_static(mk1, zh1, sh1)
    st1 = [torch.zeros((ny, nx), dtype=torch.float32, device=dev) for _ in range(4)]
    fl1 = torch.zeros((ny, nx), dtype=torch.uint8, device=dev)
    i1 = plan1.info()
    print("global stats 1-rank:", i1["mean"], i1["std2inv"], i1["r"], " chained:", stats)
STEPS = list(range(1, 17))  # 15 is a 6-hourly step at dt = 1440 s
sb_hist, inputs = [], []
for step in STEPS:
    theta, p, u, v = synth.step_fields(cfg, step, land, z, lat, seed=3)
    inputs.append((t(theta[r0:r1]), t(p[:, r0:r1]), t(u[:, r0:r1]), t(v[:, r0:r1])))
    th = ex.swap_bounds(haloed(t(theta[r0:r1]), nyl))
    plan.step(cfg.timestep, step, t(p[:, r0:r1]), t(u[:, r0:r1]), t(v[:, r0:r1]), th, *st, flags=fl)
    plan.sync()
    sb_hist.append(st[3].clone())
    pl = peer.plane(0)
    pl.fill_(7.0e30)                               # halos hold junk: they must not be read
    pl[h:h + nyl, h:h + nx] = t(theta[r0:r1])
    torch.cuda.synchronize(); peer.barrier(); torch.cuda.synchronize()
    plan_p.set_peer_theta(*peer.peers(0))
    plan_p.step(cfg.timestep, step, t(p[:, r0:r1]), t(u[:, r0:r1]), t(v[:, r0:r1]), pl, *st_p)
    plan_p.sync()
    torch.cuda.synchronize(); peer.barrier(); torch.cuda.synchronize()
    for s_a, s_b, nm in zip(st, st_p, ("ws", "wd", "thc", "sb")):
        a_, b_ = s_a.cpu().numpy(), s_b.cpu().numpy()
        same = (a_.view(np.uint32) == b_.view(np.uint32)) | (np.isnan(a_) & np.isnan(b_))
        if not same.all():
            print("rank", rank, "step", step, nm, "PEER path DIFFERENT (%d cells)" % (~same).sum())
            ok = False
    outs = [torch.empty((world, nyl, nx), dtype=torch.float32, device=dev) for _ in range(4)]
    for o, s_ in zip(outs, st):
        dist.all_gather_into_tensor(o, s_.contiguous())
    if rank == 0:
        th1 = ex1.swap_bounds(haloed(t(theta), ny))
        plan1.step(cfg.timestep, step, t(p), t(u), t(v), th1, *st1, flags=fl1)
        plan1.sync()
        for o, s1, nm in zip(outs, st1, ("ws", "wd", "thc", "sb")):
            a, b = o.reshape(ny, nx).cpu().numpy(), s1.cpu().numpy()
            same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
            print("step", step, nm, "identical" if same.all() else "DIFFERENT (%d cells)" % (~same).sum())
            ok &= bool(same.all())
# third plan: peer-memory halo reads with the post / wait words handled inside the window kernel (no
# barrier launch); all 16 steps are enqueued back to back without a host sync
peer_f = halo.PeerHalo(len(STEPS), nyl, nx, h, rank, world, dev)
for i, (th_i, _, _, _) in enumerate(inputs):
    pl = peer_f.plane(i)
    pl.fill_(7.0e30)
    pl[h:h + nyl, h:h + nx] = th_i
torch.cuda.synchronize(); dist.barrier()
plan_f = generic.Plan(SBD_VARIANT_GENERIC, nx, nyl, nz, h)
plan_f.set_halo_mode(ex.halo_mode())
plan_f.set_static(mk, zh, sh, stats=stats)
plan_f.set_peer_sync(*peer_f.sync_words())

st_f = [torch.zeros((nyl, nx), dtype=torch.float32, device=dev) for _ in range(3)]
sb_f = [torch.zeros((nyl, nx), dtype=torch.float32, device=dev) for _ in STEPS]
torch.cuda.synchronize()
for i, step in enumerate(STEPS):
    _, p_i, u_i, v_i = inputs[i]
    plan_f.set_peer_theta(*peer_f.peers(i))
    plan_f.step(cfg.timestep, step, p_i, u_i, v_i, peer_f.plane(i), *st_f, sb_f[i])
plan_f.sync()
torch.cuda.synchronize()
if peer_f.timed_out():
    print("rank", rank, "peer flag wait TIMED OUT"); ok = False
def same_bits(x, y):
    a_, b_ = x.cpu().numpy(), y.cpu().numpy()
    return bool(((a_.view(np.uint32) == b_.view(np.uint32)) | (np.isnan(a_) & np.isnan(b_))).all())
for i, step in enumerate(STEPS):
    if not same_bits(sb_hist[i], sb_f[i]):
        print("rank", rank, "step", step, "sb_con FLAG-SYNC path DIFFERENT"); ok = False
for s_a, s_b, nm in zip(st[:3], st_f, ("ws", "wd", "thc")):
    if not same_bits(s_a, s_b):
        print("rank", rank, nm, "final state FLAG-SYNC path DIFFERENT"); ok = False
dist.barrier()
okt = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(okt, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTIGPU_CHECK", "PASS" if int(okt.item()) else "FAIL", "world", world, "(NCCL exchange, peer-memory halo reads with barrier, and with in-kernel flags, vs 1 rank)")
dist.barrier()
dist.destroy_process_group()
