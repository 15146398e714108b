#!/usr/bin/env python
"""bench.py -- grid-cells/s per timestep of the sea-breeze trigger (BASELINE.json metric) on N B200s.

Workloads (BASELINE.json configs; --workload picks one, the default follows --gpus):

  c2  (default at --gpus 1)  global N512-class grid 1024 x 769 x 70, halo 6, generic seabreeze_diag
      (generic/sea_breeze_diag.f90:55-271), one timestep per step; HBM-bound.  At N > 1 (only when asked for with
      --workload c2): weak scaling, one c2-sized latitude block per rank.
  c3  (default at --gpus N > 1)  regional 2.5 km grid 4000 x 3000 x 70, halo 54, generic seabreeze_diag: ONE
      coherent grid split into N latitude row blocks -- STRONG scaling.  theta's halo rows are read from the
      neighbour ranks' planes over NVLink (symmetric memory) with the per-step handshake inside the window
      kernel.  The line also carries `strong_scaling_base`: the same grid on rank 0 alone, measured in the same
      run, so t_1 / (N t_N) can be formed from one record.
  c4  global 0.1 degree grid 3600 x 1800, 37 levels, 24 hourly records, python variant diag
      (python_wrapper/seabreezediag/seabreeze_diag_python.f90:49-285); a step is one record; N > 1: overlapping row
      blocks, no exchange (halo.PyRowBlock's scheme, device-resident).

Every line:
  value      whole-job cells/s, inputs resident in HBM, K steps back to back, CUDA events on the launching
             stream, max over ranks.
  parity_check  after the timed region every rank runs 3 steps from a zero state on the coherent global grid and
             rank 0 compares all rows, bit for bit, with its own 1-rank plan: "pass" or the process exits non-zero.
  roofline   the dominant kernel of that workload (c2: level scan, HBM; c3: window kernel, FP32 pipe; c4: window
             kernel, HBM) timed alone with CUDA events by the library.
  e2e        the same step through the public host-buffer API inside the timed region (H2D + kernels + D2H);
             c2 also has e2e_dropin: the reference's own call generic.seabreeze_diag(...) with plain numpy arrays.
  cpu_baseline  (N = 1, rank 0) the oracle -- C++ restatement of the Fortran, gfortran does not exist here -- in
             literal mode on all host threads, on a bounded sample of the same workload.
At N = 1 the default line adds `other_configs`: short c3 and c4 runs on one GPU.  At N > 1 it adds c4 on N ranks.

--impl reference: the oracle on the host cores for the same workload (rank 0 only).
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "grid_cells_per_sec_per_timestep"
UNIT = "cells/s"
NSETS = int(os.environ.get("BENCH_NSETS", "4"))  # c2: rotating input sets, 4 x 661 MB of p,u,v >> 126 MB L2


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


class Ctx:
    """rank / device / process group of this bench process"""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        from seabreeze_param_b200._capi import load
        self.lib = load()
        assert self.lib.sbd_device_count() >= 1, "no CUDA device: there is no CPU fallback"
        self.stream = torch.cuda.Stream(device=self.dev)
        torch.cuda.synchronize()
        torch.cuda.set_stream(self.stream)

    def t(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        tt = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(tt, op=self.dist.ReduceOp.MAX)
        return float(tt.item())

    def sum_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        tt = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(tt)
        return float(tt.item())

    def all_ok(self, ok):
        if self.world == 1:
            return bool(ok)
        tt = self.torch.tensor([1 if ok else 0], device=self.dev)
        self.dist.all_reduce(tt, op=self.dist.ReduceOp.MIN)
        return bool(int(tt.item()))


def same_bits(a, b):
    """bit-for-bit equality of two float32 / uint8 tensors (NaN == NaN)"""
    import torch
    if a.dtype == torch.uint8:
        return bool(torch.equal(a, b))
    ai, bi = a.contiguous().view(torch.int32), b.contiguous().view(torch.int32)
    return bool(((ai == bi) | (torch.isnan(a) & torch.isnan(b))).all().item())


def timed_steps(ctx, one_step, k0, warmup, steps, sample_clocks=True):
    """warm-up, then K steps between two events on the launching stream; returns (ms total max over ranks, k, clocks)"""
    torch = ctx.torch
    k = k0
    for _ in range(max(warmup, 3)):
        one_step(k); k += 1
    ctx.barrier()
    sampler = ClockSampler(ctx.local)
    if ctx.rank == 0 and sample_clocks:
        sampler.start()
        time.sleep(0.25)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    ev0.record(ctx.stream)
    for _ in range(steps):
        one_step(k); k += 1
    ev1.record(ctx.stream)
    ctx.barrier()
    total_ms = ctx.max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop() if (ctx.rank == 0 and sample_clocks) else None
    return total_ms, k, clocks


def balanced_rows(cost_per_row, world, align=32, min_rows=1):
    """Row-block boundaries [r_0 = 0, r_1, ..., r_world = ny] that equalise the summed cost per block; interior
    boundaries sit on multiples of `align` (the window kernels' tile height: no partial tiles at block edges)."""
    ny = len(cost_per_row)
    pre = np.concatenate([[0.0], np.cumsum(np.asarray(cost_per_row, np.float64))])
    bounds = [0]
    for k in range(1, world):
        r = int(np.searchsorted(pre, pre[-1] * k / world))
        r = int(round(r / align)) * align
        r = max(r, bounds[-1] + max(min_rows, align))
        r = min(r, ny - (world - k) * max(min_rows, align))
        bounds.append(r)
    bounds.append(ny)
    return bounds


def window_cost_rows(nn, per_cell):
    """static cost model of a latitude row: window elements of its coastal-zone cells + a per-cell constant"""
    nn = nn.astype(np.float64)
    return ((2.0 * nn + 1.0) ** 2 * (nn > 0)).sum(axis=1) + per_cell * nn.shape[1]


# ======================================================================================================
# deterministic device-side generators (a function of GLOBAL position only: every decomposition sees the same grid)
# ======================================================================================================
def hash_noise(torch, idx, salt):
    """uniform [0,1) pseudo-noise of an int64 index tensor (float64 sine hash; same bits on every rank)"""
    x = torch.sin(idx.to(torch.float64) * 12.9898 + salt * 78.233) * 43758.5453
    return (x - torch.floor(x)).to(torch.float32)


def gen_uv(ctx, nz, ny, nx, r0, r1, s, kind, chunk=8):
    """(nz, r1-r0, nx) float32 wind component of input set s for global rows [r0, r1)"""
    torch = ctx.torch
    out = torch.empty((nz, r1 - r0, nx), dtype=torch.float32, device=ctx.dev)
    yy = torch.arange(r0, r1, device=ctx.dev, dtype=torch.int64)[None, :, None]
    xx = torch.arange(nx, device=ctx.dev, dtype=torch.int64)[None, None, :]
    base = (6.0 * torch.sin(2 * math.pi * (3.0 * xx.double() / nx + 0.11 * s + 0.5 * kind))
            * torch.cos(2 * math.pi * (2.0 * yy.double() / ny + 0.07 * s))).to(torch.float32)
    for k0 in range(0, nz, chunk):
        k1 = min(nz, k0 + chunk)
        kk = torch.arange(k0, k1, device=ctx.dev, dtype=torch.int64)[:, None, None]
        idx = (kk * ny + yy) * nx + xx
        scale = (0.6 + 0.8 * kk.to(torch.float32) / max(nz - 1, 1))
        out[k0:k1] = base * scale + 2.0 * (hash_noise(torch, idx, 1000.0 * kind + s) - 0.5)
    return out


def gen_p3d(ctx, z_rows, nz):
    """generic variant: p(k) = 101325 exp(-z/8000) exp(-k ln(100)/(nz-1)) -- monotone, 700 hPa near k ~ 6"""
    torch = ctx.torch
    k = torch.arange(nz, device=ctx.dev, dtype=torch.float32)[:, None, None]
    return (101325.0 * torch.exp(-z_rows / 8000.0)[None] * torch.exp(-k * math.log(100.0) / max(nz - 1, 1))).contiguous()


def theta_host(cfg, s, land, lat, hour):
    """theta of input set s on the host (2-D only; synth.step_fields' formula without the 3-D fields)"""
    rng = np.random.Generator(np.random.PCG64(1000 + s))
    latr = np.deg2rad(lat.astype(np.float64))[:, None]
    th = (300.0 - 30.0 * np.sin(latr) ** 2 + 3.0 * land * math.cos(2 * math.pi * hour / 24.0)
          + 0.5 * rng.standard_normal(land.shape))
    return th.astype(np.float32)


# ======================================================================================================
# c2: N512-class grid, generic variant
# ======================================================================================================
def build_case(cfg, rank, seed0=0):
    """Static fields + NSETS step inputs for one latitude block (numpy, host)."""
    from seabreeze_param_b200 import synth
    land, ice, lon, lat, z, sig = synth.static_fields(cfg, seed=seed0 + 17 * rank)
    steps = []
    for s in range(NSETS):
        theta, p, u, v = synth.step_fields(cfg, s + 1, land, z, lat, seed=seed0 + 17 * rank)
        steps.append((theta, p, u, v))
    return land, ice, lon, lat, z, sig, steps


def algorithmic_bytes(n_cells, n_coastal, nz, six_hourly_frac=0.0):
    """B_diag(generic, step>=2) = 20 N + nC (4 nz + 24) (+4 nC on a 6-hourly step)."""
    return 20.0 * n_cells + n_coastal * (4.0 * nz + 24.0 + 4.0 * six_hourly_frac)


def bench_c2(ctx, args):
    import torch
    from seabreeze_param_b200 import generic, halo, synth
    from seabreeze_param_b200._capi import SBD_VARIANT_GENERIC
    cfg = synth.CONFIGS["c2"]
    world, rank, dev, dist = ctx.world, ctx.rank, ctx.dev, ctx.dist
    h, nx, ny, nz = cfg.h, cfg.nx, cfg.ny, cfg.nz
    land, ice, lon, lat, z, sig, steps = build_case(cfg, rank)
    t = ctx.t
    mask_h = torch.zeros((ny + 2 * h, nx + 2 * h), dtype=torch.float32, device=dev)
    generic.get_edges(mask_h, t(ice), t(land), h)
    cdist = torch.empty((ny, nx), dtype=torch.float32, device=dev)
    generic.get_dist(mask_h, t(land), t(lon), t(lat), 180.0, cdist, h)

    def haloed(a2d):
        f = torch.zeros((ny + 2 * h, nx + 2 * h), dtype=torch.float32, device=dev)
        f[h:h + ny, h:h + nx] = a2d
        return f

    ex = halo.RowBlockHalo(nx, ny, h, rank, world, dev)
    mk = ex.swap_bounds(haloed(cdist))
    zh = ex.swap_bounds(haloed(t(z)))
    sh = ex.swap_bounds(haloed(t(sig)))
    peer = None
    if world > 1:
        peer = halo.PeerHalo(NSETS, ny, nx, h, rank, world, dev)
    dsteps = []
    for i, (theta, p, u, v) in enumerate(steps):
        if peer is not None:
            th = peer.plane(i)
            th[h:h + ny, h:h + nx] = t(theta)
        else:
            th = haloed(t(theta))
        dsteps.append((th, t(p), t(u), t(v)))
    ctx.barrier()
    state = [torch.zeros((ny, nx), dtype=torch.float32, device=dev) for _ in range(4)]
    flags = torch.zeros((ny, nx), dtype=torch.uint8, device=dev)
    plan = generic.Plan(SBD_VARIANT_GENERIC, nx, ny, nz, h)
    plan.set_stream(ctx.stream.cuda_stream)
    plan.set_halo_mode(ex.halo_mode())
    stats = ex.global_sigma_stats(sh) if world > 1 else None
    plan.set_static(mk, zh, sh, stats=stats)       # first call: module load, cub temp storage, ...
    torch.cuda.synchronize()
    t_static0 = time.perf_counter()
    plan.set_static(mk, zh, sh, stats=stats)       # warm: what a host pays when its static fields change
    torch.cuda.synchronize()
    static_ms = 1e3 * (time.perf_counter() - t_static0)
    info = plan.info()
    n_cells = nx * ny
    n_coastal = info["n_coastal"]
    if peer is not None:
        plan.set_peer_sync(*peer.sync_words())

    def one_step(k, st=state, fl=flags):
        th, p, u, v = dsteps[k % NSETS]
        if peer is not None:
            plan.set_peer_theta(*peer.peers(k % NSETS))
        plan.step(cfg.timestep, k + 2, p, u, v, th, *st, flags=fl)

    l0 = ctx.lib.sbd_launch_count()
    total_ms, k, clocks = timed_steps(ctx, one_step, 0, args.warmup, args.steps)
    launches = (ctx.lib.sbd_launch_count() - l0) * args.steps // (args.steps + max(args.warmup, 3))
    plan.set_timing(True)
    win_ms, kern_ms = [], []
    for i in range(min(args.steps, 20)):
        one_step(k)
        a_ms, b_ms, _ = plan.last_timing()
        win_ms.append(a_ms); kern_ms.append(b_ms)
        k += 1
    plan.set_timing(False)
    if peer is not None and peer.timed_out():
        raise SystemExit("bench: a window kernel timed out waiting for a neighbour rank's halo rows")
    ms_per_step = total_ms / args.steps
    value = world * n_cells * args.steps / (total_ms * 1e-3)
    n_coastal_all = int(ctx.sum_over_ranks(n_coastal))

    # ---- parity check: N = 1: the plan path against the reference's own one-shot call, 3 steps from zero state;
    #      N > 1 (weak scaling, independent blocks per rank): the same, per rank
    parity = "pass"
    st_a = [torch.zeros((ny, nx), dtype=torch.float32, device=dev) for _ in range(4)]
    st_b = [torch.zeros((ny, nx), dtype=torch.float32, device=dev) for _ in range(4)]
    fl_a = torch.zeros((ny, nx), dtype=torch.uint8, device=dev); fl_b = torch.zeros_like(fl_a)
    if peer is None:
        for s in (1, 2, 15):
            th, p, u, v = dsteps[s % NSETS]
            plan.step(cfg.timestep, s, p, u, v, th, *st_a, flags=fl_a)
            thh = ex.swap_bounds(th.clone())
            generic.seabreeze_diag(cfg.timestep, s, p, u, v, thh, mk, zh, sh, *st_b, flags=fl_b)
            plan.sync()
            if not all(same_bits(a, b) for a, b in zip(st_a + [fl_a], st_b + [fl_b])):
                parity = "FAIL: plan step != one-shot seabreeze_diag at step %d" % s
    if not ctx.all_ok(parity == "pass"):
        parity = parity if parity != "pass" else "FAIL on another rank"

    # ---- e2e: host buffers through the public API ----
    e2e = e2e_dropin = None
    if not args.no_e2e:
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
        hsteps = []
        for i in range(2):
            theta, p, u, v = steps[i]
            hsteps.append((pin(synth.pad_halo(theta, h)), pin(p), pin(u), pin(v)))
        hstate = [pin(np.zeros((ny, nx), np.float32)) for _ in range(4)]
        hflags = torch.zeros((ny, nx), dtype=torch.uint8).pin_memory().numpy()
        if peer is not None:
            plan.set_peer_theta(None, 0, None)
            plan.set_peer_sync(None, None, None)
        ne = max(3, min(args.steps, 10))
        for i in range(2):
            th_h, p, u, v = hsteps[i % 2]
            plan.step(cfg.timestep, i + 2, p, u, v, th_h, *hstate, flags=hflags)
        ctx.barrier()
        t0 = time.perf_counter()
        for i in range(ne):
            th_h, p, u, v = hsteps[i % 2]
            plan.step(cfg.timestep, i + 4, p, u, v, th_h, *hstate, flags=hflags)  # blocking: outputs are on the host
        torch.cuda.synchronize()
        dt = ctx.max_over_ranks(time.perf_counter() - t0)
        nq, nqn, ysp = info["n_quads"], info["n_quads_north"], info["host_split_row"]
        h2d = (4 * ysp * nx * nz + 16 * (nq - nqn) * nz + 16 * nq * 2
               + 4 * ((nx + 2 * h) * (ny + 2 * h) + 3 * n_cells))
        d2h = 4 * (4 * n_cells) + n_cells
        e2e = {"value": world * n_cells * ne / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": ne, "ms_per_step": 1e3 * dt / ne,
               "api": "Plan.step with numpy arrays in pinned host memory (blocking call); p: DMA of the rows that "
                      "hold coastal-zone cells (up to the split row) + zero-copy gather of the columns south of it",
               "host_input_bytes_per_step": 4 * (3 * n_cells * nz + (nx + 2 * h) * (ny + 2 * h) + 3 * n_cells)}
        if world == 1:
            # the reference's own call with plain numpy arrays (np.array copies: pageable, never seen by CUDA before):
            # generic.seabreeze_diag -> sbd_seabreeze_diag_f32.  First calls page-lock the arrays (auto_register) and
            # build the cached plan; the timed calls are what a host model pays per timestep after that.
            # (halo mode 0 of the one-shot call reads theta's halo memory: fill it as the host's swap_bounds would)
            psteps = [(generic.swap_bounds(np.array(hs[0]), h),) + tuple(np.array(a) for a in hs[1:]) for hs in hsteps]
            pstate = [np.zeros((ny, nx), np.float32) for _ in range(4)]
            pflags = np.zeros((ny, nx), np.uint8)
            mk_n, zh_n, sh_n = (np.array(x.cpu().numpy()) for x in (mk, zh, sh))
            t_first0 = time.perf_counter()
            for i in range(2):
                th_h, p, u, v = psteps[i % 2]
                generic.seabreeze_diag(cfg.timestep, i + 2, p, u, v, th_h, mk_n, zh_n, sh_n, *pstate, flags=pflags)
            first_ms = 1e3 * (time.perf_counter() - t_first0) / 2
            t0 = time.perf_counter()
            for i in range(ne):
                th_h, p, u, v = psteps[i % 2]
                generic.seabreeze_diag(cfg.timestep, i + 4, p, u, v, th_h, mk_n, zh_n, sh_n, *pstate, flags=pflags)
            dt = time.perf_counter() - t0
            e2e_dropin = {"value": n_cells * ne / dt, "unit": UNIT, "steps": ne, "ms_per_step": 1e3 * dt / ne,
                          "h2d_bytes_per_step": h2d + 3 * 4 * (nx + 2 * h) * (ny + 2 * h), "d2h_bytes_per_step": d2h,
                          "first_calls_ms": first_ms,
                          "api": "generic.seabreeze_diag(timestep, n, p,u,v, theta, mask, z, sigma, ws, wd, thc, sb_con) = "
                                 "sbd_seabreeze_diag_f32, the call of generic/dummy_model.f90:52-54, plain np.array inputs; "
                                 "static planes uploaded + checksummed every call (plan cache), p,u,v page-locked on first "
                                 "sight (first_calls_ms) and gathered in place afterwards"}
            ok = all(np.array_equal(a.view(np.uint32), b.view(np.uint32)) for a, b in zip(pstate, hstate))
            if not ok:
                parity = "FAIL: one-shot host call != plan host call"
            ctx.lib.sbd_plan_cache_clear()

    if rank != 0:
        return None
    peak, peak_src = measured_peak()
    six_frac = sum(1 for i in range(args.steps) if ((k - args.steps + i + 2) * cfg.timestep) % 21600 == 0) / args.steps
    abytes = n_coastal * (4.0 * nz + 28.0 + 4.0 * six_frac)
    abytes_step = algorithmic_bytes(n_cells, n_coastal, nz, six_frac)
    kms = statistics.mean(kern_ms)
    achieved = abytes / (kms * 1e-3) / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "traffic_k_scan_quad_regs.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            pass
    roofline = {"bound": "hbm",
                "kernel": "sbdk::k_scan_quad_regs<5,128> (level scan over p + thresholds; timed alone.  The window kernel "
                          "sbdk::k_trigger, timed alone: %.2f us; in the timed steps the two overlap)" % (1e3 * statistics.mean(win_ms)),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": abytes, "algorithmic_bytes_per_step": abytes_step,
                "moved_bytes_frac": (traffic / (kms * 1e-3) / 1e9 / peak) if traffic else None,
                "step_achieved_gbs": abytes_step / (ms_per_step * 1e-3) / 1e9,
                "step_frac": abytes_step / (ms_per_step * 1e-3) / 1e9 / peak, "kernel_ms": kms,
                "kernel_ms_min": min(kern_ms), "kernel_share_of_step": kms / ms_per_step,
                "floor_note": "scripts/mb_scan.cu (profiles/r2_mb_scan.txt): this gather costs the same ~35 us as a dense "
                              "read of all 220 MB of p -- the cost follows the 128-byte lines touched (97 % of them)"}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu = cpu_c2(cfg, steps, [x.cpu().numpy() for x in (mk, zh, sh)], [d[0].cpu().numpy() for d in dsteps])
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "c2: " + cfg.descr + (" per GPU (latitude blocks, %d ranks)" % world if world > 1 else ""),
                   "nx": nx, "ny": ny * world, "nz": nz, "halo": h, "n_coastal_cells": n_coastal_all,
                   "api": "plan step (static fields processed once per plan: %.1f ms warm, SURVEY 8d)" % static_ms,
                   "l2": "rotating %d input sets of p,u,v (%.0f MB each) > 126 MB L2; no flush kernel in the timed region"
                         % (NSETS, 3 * 4 * n_cells * nz / 1e6),
                   "parallelism": ("row-block x%d, theta halo rows read from the neighbours' planes over NVLink (symmetric "
                                   "memory), post/wait flags inside the window kernel" % world) if world > 1 else "single GPU"},
        "parity_check": parity, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_dropin": e2e_dropin,
        "gpu_launches": int(launches), "clocks": clocks, "static_setup_ms": static_ms,
    }
    return line


def cpu_c2(cfg, steps, statics, thetas, reps=3):
    from oracle import oracle as O
    O.set_math(0)
    nthreads = O.max_threads()
    O.set_threads(nthreads)
    mk_h, zh_h, sh_h = statics
    st = [np.zeros((cfg.ny, cfg.nx), np.float32) for _ in range(4)]
    tb = []
    for i in range(reps):
        theta, p, u, v = steps[i % len(steps)]
        t0 = time.perf_counter()
        O.seabreeze_diag(cfg.timestep, i + 1, p, u, v, thetas[i % len(thetas)], mk_h, zh_h, sh_h, *st, cfg.h, literal=True)
        tb.append(time.perf_counter() - t0)
    O.set_math(1)
    O.set_threads(1)
    return {"value": cfg.nx * cfg.ny / min(tb), "unit": UNIT, "cores": nthreads, "kind": "port",
            "sample": "full c2 grid, best of %d steps, C++ restatement of the Fortran (gfortran unavailable), "
                      "literal loops, platform libm" % reps, "ms_per_step": 1e3 * min(tb)}


# ======================================================================================================
# c3: regional 2.5 km grid, one coherent grid over N ranks (strong scaling)
# ======================================================================================================
class C3Grid:
    """global static planes of the c3 grid on this rank's device (the static chain is cheap and one-off: every
    rank evaluates it on the whole grid and cuts its rows, so no static exchange is needed)"""

    def __init__(self, ctx, cfg):
        import torch
        from seabreeze_param_b200 import generic, synth
        self.cfg = cfg
        h, nx, ny = cfg.h, cfg.nx, cfg.ny
        dev = ctx.dev
        self.land, self.ice, self.lon, self.lat, self.z, self.sig = synth.static_fields(cfg)
        t = ctx.t
        mask_h = torch.zeros((ny + 2 * h, nx + 2 * h), dtype=torch.float32, device=dev)
        generic.get_edges(mask_h, t(self.ice), t(self.land), h)
        cdist = torch.empty((ny, nx), dtype=torch.float32, device=dev)
        generic.get_dist(mask_h, t(self.land), t(self.lon), t(self.lat), 180.0, cdist, h)
        del mask_h

        def haloed(a2d):
            f = torch.zeros((ny + 2 * h, nx + 2 * h), dtype=torch.float32, device=dev)
            f[h:h + ny, h:h + nx] = a2d
            generic.swap_bounds(f, h)
            return f
        self.haloed = haloed
        self.z_dev = t(self.z)
        self.mk, self.zh, self.sh = haloed(cdist), haloed(self.z_dev), haloed(t(self.sig))
        st = np.zeros(4, np.float32)
        generic.sigmoid(t(self.sig), stats=st)           # global sequential statistics: same bits on every rank
        self.stats = st[:3].copy()

    def theta_h(self, ctx, s):
        return self.haloed(ctx.t(theta_host(self.cfg, s, self.land, self.lat, hour=0.4 * (s + 1))))


def bench_c3(ctx, args, steps, warmup, short=False):
    import torch
    from seabreeze_param_b200 import generic, halo, synth
    from seabreeze_param_b200._capi import SBD_VARIANT_GENERIC
    cfg = synth.CONFIGS["c3"]
    world, rank, dev, dist = ctx.world, ctx.rank, ctx.dev, ctx.dist
    h, nx, ny, nz = cfg.h, cfg.nx, cfg.ny, cfg.nz
    nsets = 1 if world == 1 else 2     # one set of p,u,v is 10 GB / N per rank: far beyond L2 already
    G = C3Grid(ctx, cfg)
    n_cells = nx * ny
    # Row blocks of EQUAL WORK, not equal height: the search radius -- hence the window sums, (2nn+1)^2 adds per
    # coastal cell -- varies strongly with latitude band.  The radius plane is static, so every rank evaluates it
    # once on the whole grid (the static chain runs there anyway) and all ranks derive the same boundaries.
    # BENCH_C3_SPLIT=equal: blocks of equal height (measured at 8 GPUs: 1.04 ms/step, the slowest rank's window
    # kernel 0.97 ms against a mean of 0.5 ms).
    plan_s = generic.Plan(SBD_VARIANT_GENERIC, nx, ny, nz, h)
    plan_s.set_static(G.mk, G.zh, G.sh)
    nn_g, _, _ = plan_s.window_info()
    info_full = plan_s.info()
    plan_s.close()
    if os.environ.get("BENCH_C3_SPLIT", "work") == "equal" or world == 1:
        bounds = [round(ny * k / world) for k in range(world + 1)]
    else:
        bounds = balanced_rows(window_cost_rows(nn_g, per_cell=60.0), world, align=32, min_rows=h)
    rows_all = [bounds[k + 1] - bounds[k] for k in range(world)]
    r0, r1 = bounds[rank], bounds[rank + 1]
    nyl = r1 - r0
    nyl_max = max(rows_all)

    # ---- rank 0: the whole grid on one GPU (strong-scaling base + the reference bits of the parity check) ----
    base = None
    full_inputs = None
    if rank == 0:
        plan1 = generic.Plan(SBD_VARIANT_GENERIC, nx, ny, nz, h)
        plan1.set_stream(ctx.stream.cuda_stream)
        plan1.set_static(G.mk, G.zh, G.sh)
        th1 = G.theta_h(ctx, 0)
        p1 = gen_p3d(ctx, G.z_dev, nz)
        u1, v1 = gen_uv(ctx, nz, ny, nx, 0, ny, 0, 0), gen_uv(ctx, nz, ny, nx, 0, ny, 0, 1)
        full_inputs = (th1, p1, u1, v1)
        st1 = [torch.zeros((ny, nx), dtype=torch.float32, device=dev) for _ in range(4)]
        fl1 = torch.zeros((ny, nx), dtype=torch.uint8, device=dev)
        info1 = info_full
        if True:
            for s in range(3):
                plan1.step(cfg.timestep, s + 2, p1, u1, v1, th1, *st1, flags=fl1)
            torch.cuda.synchronize()
            nb = max(3, min(steps, 10))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ctx.stream)
            for s in range(nb):
                plan1.step(cfg.timestep, s + 5, p1, u1, v1, th1, *st1, flags=fl1)
            e1.record(ctx.stream)
            torch.cuda.synchronize()
            t1 = e0.elapsed_time(e1) / nb
            plan1.set_timing(True)
            w1, s1 = [], []
            for s in range(3):
                plan1.step(cfg.timestep, s + 20, p1, u1, v1, th1, *st1, flags=fl1)
                a_ms, b_ms, _ = plan1.last_timing()
                w1.append(a_ms); s1.append(b_ms)
            plan1.set_timing(False)
            base = {"n_gpus": 1, "ms_per_step": t1, "value": n_cells / (t1 * 1e-3), "steps": nb,
                    "window_kernel_ms": statistics.mean(w1), "scan_kernel_ms": statistics.mean(s1),
                    "nn_max": info1["nn_max"], "n_coastal_cells": info1["n_coastal"],
                    "measured": "same run, rank 0 alone, whole 4000 x 3000 x 70 grid resident on one GPU"}
    ctx.barrier()

    # ---- this rank's row block ----
    sl = slice(r0, r1 + 2 * h)                        # rows r0-h .. r1+h of the global interior = the block's halo'd plane
    mk_b, zh_b, sh_b = (x[sl].contiguous() for x in (G.mk, G.zh, G.sh))
    plan = generic.Plan(SBD_VARIANT_GENERIC, nx, nyl, nz, h)
    plan.set_stream(ctx.stream.cuda_stream)
    mode = 1 | (2 if rank == 0 else 0) | (4 if rank == world - 1 else 0)
    plan.set_halo_mode(mode)
    plan.set_static(mk_b, zh_b, sh_b, stats=G.stats)
    info = plan.info()
    peer = halo.PeerHalo(nsets, nyl_max, nx, h, rank, world, dev) if world > 1 else None   # symmetric: same shape everywhere
    ny_up = rows_all[rank - 1] if rank > 0 else 0

    def peers(i):
        up, _, dn = peer.peers(i)
        return up, ny_up, dn
    dsteps = []
    for s in range(nsets):
        thg = G.theta_h(ctx, s)
        if peer is not None:
            th = peer.plane(s)[:nyl + 2 * h]           # this rank's (nyl + 2h)-row plane inside the symmetric buffer
            th.fill_(7.0e30)                           # halo rows owned by a neighbour hold junk: they must not be read
            th[h:h + nyl, :] = thg[r0 + h:r1 + h, :]
        else:
            th = thg[sl].contiguous()
        dsteps.append((th, gen_p3d(ctx, G.z_dev[r0:r1], nz), gen_uv(ctx, nz, ny, nx, r0, r1, s, 0),
                       gen_uv(ctx, nz, ny, nx, r0, r1, s, 1)))
        del thg
    ctx.barrier()
    if peer is not None:
        plan.set_peer_sync(*peer.sync_words())
    state = [torch.zeros((nyl, nx), dtype=torch.float32, device=dev) for _ in range(4)]
    flags = torch.zeros((nyl, nx), dtype=torch.uint8, device=dev)

    def one_step(k, st=state, fl=flags, first=2):
        th, p, u, v = dsteps[k % nsets]
        if peer is not None:
            plan.set_peer_theta(*peers(k % nsets))
        plan.step(cfg.timestep, k + first, p, u, v, th, *st, flags=fl)

    l0 = ctx.lib.sbd_launch_count()
    total_ms, k, clocks = timed_steps(ctx, one_step, 0, warmup, steps, sample_clocks=not short)
    launches = (ctx.lib.sbd_launch_count() - l0) * steps // (steps + max(warmup, 3))
    plan.set_timing(True)
    win_ms, scan_ms = [], []
    for i in range(3):
        ctx.barrier()     # ranks start the step together: the window kernel's wait for the neighbours' posts stays short
        one_step(k)
        a_ms, b_ms, _ = plan.last_timing()
        win_ms.append(a_ms); scan_ms.append(b_ms)
        k += 1
    plan.set_timing(False)
    win = ctx.max_over_ranks(statistics.mean(win_ms))
    scn = ctx.max_over_ranks(statistics.mean(scan_ms))
    win_all = [statistics.mean(win_ms)]
    if world > 1:
        wt = torch.tensor([statistics.mean(win_ms)], dtype=torch.float64, device=dev)
        wl = [torch.zeros_like(wt) for _ in range(world)]
        dist.all_gather(wl, wt)
        win_all = [float(x.item()) for x in wl]
    ms_per_step = total_ms / steps
    value = n_cells * steps / (total_ms * 1e-3)
    n_coastal_all = int(ctx.sum_over_ranks(info["n_coastal"]))

    # ---- parity check: 3 steps from a zero state with input set 0, every rank's rows vs rank 0's 1-rank plan ----
    parity = "pass"
    stp = [torch.zeros((nyl, nx), dtype=torch.float32, device=dev) for _ in range(4)]
    flp = torch.zeros((nyl, nx), dtype=torch.uint8, device=dev)
    if rank == 0:
        stf = [torch.zeros((ny, nx), dtype=torch.float32, device=dev) for _ in range(4)]
        flf = torch.zeros((ny, nx), dtype=torch.uint8, device=dev)
    for s in (1, 2, 15):
        th, p, u, v = dsteps[0]
        ctx.barrier()                                   # set 0's theta plane of every rank is in place and unread
        if peer is not None:
            plan.set_peer_theta(*peers(0))
        plan.step(cfg.timestep, s, p, u, v, th, *stp, flags=flp)
        plan.sync()
        if rank == 0:
            th1, p1, u1, v1 = full_inputs
            plan1.step(cfg.timestep, s, p1, u1, v1, th1, *stf, flags=flf)
            plan1.sync()
        for j, own in enumerate(stp + [flp]):
            if world > 1:
                if rank == 0:
                    parts = [own] + [torch.empty((rows_all[r], nx), dtype=own.dtype, device=dev) for r in range(1, world)]
                    for r in range(1, world):
                        dist.recv(parts[r], src=r)
                    got = torch.cat(parts, 0)
                else:
                    dist.send(own.contiguous(), dst=0)
                    got = None
            else:
                got = own
            if rank == 0:
                ref = (stf + [flf])[j]
                if not same_bits(got, ref):
                    parity = "FAIL: %s differs from the 1-rank plan at step %d" % (("windspeed", "winddir", "thc", "sb_con", "flags")[j], s)
    if peer is not None and peer.timed_out():
        parity = "FAIL: peer flag wait timed out"
    if not ctx.all_ok(parity == "pass") and parity == "pass":
        parity = "FAIL on another rank"
    n_trig = int(((flp & 2) > 0).sum().item()) if True else 0
    n_trig = int(ctx.sum_over_ranks(n_trig))

    # ---- e2e: this rank's block through the host-buffer API ----
    e2e = None
    if not args.no_e2e and not short:
        if peer is not None:
            plan.set_peer_theta(None, 0, None)
            plan.set_peer_sync(None, None, None)
        thg = G.theta_h(ctx, 0)
        th_host = np.array(thg[sl].cpu().numpy())      # halo rows valid on the host side: no peer reads
        del thg
        hp, hu, hv = (np.array(x.cpu().numpy()) for x in dsteps[0][1:])
        hstate = [np.zeros((nyl, nx), np.float32) for _ in range(4)]
        hflags = np.zeros((nyl, nx), np.uint8)
        from seabreeze_param_b200 import _capi
        for a in (hp, hu, hv, th_host, hflags, *hstate):
            _capi.auto_register(a)
        for i in range(2):
            plan.step(cfg.timestep, i + 2, hp, hu, hv, th_host, *hstate, flags=hflags)
        ctx.barrier()
        ne = 3
        t0 = time.perf_counter()
        for i in range(ne):
            plan.step(cfg.timestep, i + 4, hp, hu, hv, th_host, *hstate, flags=hflags)
        torch.cuda.synchronize()
        dt = ctx.max_over_ranks(time.perf_counter() - t0)
        nq, nqn, ysp = info["n_quads"], info["n_quads_north"], info["host_split_row"]
        h2d = world * (4 * ysp * nx * nz + 16 * (nq - nqn) * nz + 16 * nq * 2 + 4 * ((nx + 2 * h) * (nyl + 2 * h) + 3 * nyl * nx))
        e2e = {"value": n_cells * ne / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": 4 * 4 * n_cells + n_cells, "steps": ne, "ms_per_step": 1e3 * dt / ne,
               "api": "Plan.step per rank with page-locked numpy arrays of its row block (blocking); p: DMA of the "
                      "northern rows + zero-copy gather of the rest; whole-job bytes"}
        del hp, hu, hv

    if rank != 0:
        return None
    peak, peak_src = measured_peak()
    # window kernel: sum over coastal cells of (2nn+1)^2 window elements, each advancing two fp32 running sums
    elem_cells = float(((2.0 * nn_g.astype(np.float64) + 1.0) ** 2 * (nn_g > 0)).sum())
    fp32_peak = 148 * 128 * 1.965e9            # scalar FADD lanes / s (scripts/mb_pipes.cu: 120-123 of 128 per clk per SM)
    adds = 2.0 * elem_cells
    abytes_step = algorithmic_bytes(n_cells, n_coastal_all, nz)
    roofline = {"bound": "fp32-pipe",
                "kernel": "sbdk::k_win8 (window sums, %d ranks; timed alone, max over ranks: %.3f ms; level scan alone %.3f ms)" % (world, win, scn),
                "achieved": adds / (win * 1e-3) / 1e12, "peak": fp32_peak / 1e12, "unit": "TFLOP/s",
                "frac": adds / (win * 1e-3) / (fp32_peak * world), "traffic": None,
                "window_element_cells_per_step": elem_cells, "fp32_adds_per_step": adds,
                "peak_source": "148 SMs x 128 FP32 lanes x 1.965 GHz (add rate measured with scripts/mb_pipes.cu: 94-96 % of it)",
                "kernel_ms": win, "kernel_ms_per_rank": win_all, "kernel_share_of_step": win / ms_per_step,
                "hbm_algorithmic_bytes_per_step": abytes_step,
                "hbm_step_frac": abytes_step / (ms_per_step * 1e-3) / 1e9 / (peak * world), "hbm_peak": peak, "hbm_peak_source": peak_src}
    cpu = None
    if not args.no_cpu_baseline and world == 1 and not short:
        cpu = cpu_c3(cfg, G, full_inputs)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": max(warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "c3: " + cfg.descr + ", ONE grid over %d rank(s)" % world,
                   "nx": nx, "ny": ny, "nz": nz, "halo": h, "n_coastal_cells": n_coastal_all, "nn_max": info["nn_max"],
                   "rows_per_rank": rows_all, "row_split": "equal work (static radius plane), boundaries on multiples of 32 rows"
                   if world > 1 else "one block", "trigger_cells_last_parity_step": n_trig,
                   "api": "plan step per rank (static chain once per plan, evaluated on the whole grid by every rank)",
                   "l2": "%d input set(s) of p,u,v per rank, %.2f GB each >> 126 MB L2" % (nsets, 3 * 4 * nyl * nx * nz / 1e9),
                   "parallelism": ("latitude row blocks x%d; theta halo rows (54 per side) read from the neighbours' planes "
                                   "over NVLink inside the window kernel, post/wait flags in symmetric memory" % world)
                   if world > 1 else "single GPU"},
        "parity_check": parity, "strong_scaling_base": base, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": int(launches), "clocks": clocks,
    }
    return line


def cpu_c3(cfg, G, full_inputs, rows=48):
    """oracle, literal growing-window loops, on a block of `rows` rows cut out of the grid's middle (halos from the
    full planes; the block's own sigma statistics)"""
    from oracle import oracle as O
    h, nx = cfg.h, cfg.nx
    a = cfg.ny // 2
    sl = slice(a, a + rows + 2 * h)
    mk, zh, sh = (np.ascontiguousarray(x[sl].cpu().numpy()) for x in (G.mk, G.zh, G.sh))
    th1, p1, u1, v1 = full_inputs
    th = np.ascontiguousarray(th1[sl].cpu().numpy())
    p, u, v = (np.ascontiguousarray(x[:, a:a + rows].cpu().numpy()) for x in (p1, u1, v1))
    O.set_math(0)
    nthreads = O.max_threads()
    O.set_threads(nthreads)
    st = [np.zeros((rows, nx), np.float32) for _ in range(4)]
    tb = []
    for i in range(2):
        t0 = time.perf_counter()
        O.seabreeze_diag(cfg.timestep, i + 1, p, u, v, th, mk, zh, sh, *st, h, literal=True)
        tb.append(time.perf_counter() - t0)
    O.set_math(1)
    O.set_threads(1)
    return {"value": rows * nx / min(tb), "unit": UNIT, "cores": nthreads, "kind": "port",
            "sample": "%d-row block (%d x %d x %d cells, halo %d) of the c3 grid, best of 2 steps, C++ restatement of the "
                      "Fortran (gfortran unavailable), literal growing-window loops, platform libm" % (rows, nx, rows, cfg.nz, h),
            "ms_per_sample_step": 1e3 * min(tb)}


# ======================================================================================================
# c4: global 0.1 degree grid, python variant, 24 hourly records
# ======================================================================================================
def bench_c4(ctx, args, steps, warmup, short=False):
    import ctypes as C
    import torch
    from seabreeze_param_b200 import generic, synth, _capi
    from seabreeze_param_b200._capi import SBD_VARIANT_PYTHON
    cfg = synth.CONFIGS["c4"]
    world, rank, dev, dist = ctx.world, ctx.rank, ctx.dev, ctx.dist
    nx, ny, nps = cfg.nx, cfg.ny, cfg.nz
    nrec = 6 if short else 24
    lib, q, t = ctx.lib, _capi.ptr, ctx.t
    land, ice, lon, lat, z, sig = synth.static_fields(cfg)
    p = t(synth.pressure_1d(nps))
    dl, di, dz, ds = t(land), t(ice), t(z), t(sig)
    coast = torch.empty((ny, nx), dtype=torch.float32, device=dev)
    cdist = torch.empty_like(coast)
    kk = C.c_int(0)
    _capi.check(lib.sbd_py_get_edges_f32(q(dl), q(di), nx, ny, q(coast)))
    dlon, dlat = t(lon), t(lat)                        # named: a temporary would be freed (and its block reused) before the call reads it
    _capi.check(lib.sbd_py_get_dist_f32(q(coast), q(dl), q(dlon), q(dlat), nx, ny, 180.0, q(cdist), C.byref(kk)))
    st4 = np.zeros(4, np.float32)
    generic.sigmoid(ds, stats=st4)                     # global sequential statistics of std (P:291-309)
    stats = st4[:3].copy()
    overlap = 24                                        # >= the largest search radius (checked below)
    # row blocks of equal work (static radius plane + a per-cell constant for the pointwise part), see bench_c3
    if world > 1 and os.environ.get("BENCH_C4_SPLIT", "work") != "equal":
        plan_s = generic.Plan(SBD_VARIANT_PYTHON, nx, ny, nps)
        plan_s.set_static(cdist, dz, ds)
        nn_g, _, _ = plan_s.window_info()
        plan_s.close()
        bounds = balanced_rows(window_cost_rows(nn_g, per_cell=150.0), world, align=32, min_rows=overlap)
    else:
        bounds = [round(ny * k / world) for k in range(world + 1)]
    r0, r1 = bounds[rank], bounds[rank + 1]
    a, b = max(0, r0 - overlap), min(ny, r1 + overlap)
    nb = b - a
    plan = generic.Plan(SBD_VARIANT_PYTHON, nx, nb, nps)
    plan.set_stream(ctx.stream.cuda_stream)
    plan.set_row_window(r0 - a, r1 - a)
    plan.set_static(cdist[a:b].contiguous(), dz[a:b].contiguous(), ds[a:b].contiguous(), stats=stats)
    info = plan.info()
    if info["nn_max"] > overlap:
        raise SystemExit("c4: overlap %d < largest search radius %d" % (overlap, info["nn_max"]))

    def record(rows0, rows1, r):
        """theta (rows, nx), u, v (nps, rows, nx) of hourly record r for global rows [rows0, rows1)"""
        th = ctx.t(theta_host(cfg, 100 + r, land, lat, hour=float(r))[rows0:rows1])
        return th, gen_uv(ctx, nps, ny, nx, rows0, rows1, 100 + r, 0), gen_uv(ctx, nps, ny, nx, rows0, rows1, 100 + r, 1)

    recs = [record(a, b, r) for r in range(nrec)]
    outs = [torch.zeros((4, nb, nx), dtype=torch.float32, device=dev) for _ in range(2)]
    zero = torch.zeros((nb, nx), dtype=torch.float32, device=dev)
    flags = torch.zeros((nb, nx), dtype=torch.uint8, device=dev)
    ctx.barrier()

    def one_step(k, outs=outs, fl=flags, first=2):
        th, u, v = recs[k % nrec]
        prev = outs[(k + 1) % 2]
        plan.step_py(k + first, cfg.timestep, p, u, v, th, prev[2], prev[3], outs[k % 2], flags=fl)

    l0 = ctx.lib.sbd_launch_count()
    total_ms, k, clocks = timed_steps(ctx, one_step, 0, warmup, steps, sample_clocks=not short)
    launches = (ctx.lib.sbd_launch_count() - l0) * steps // (steps + max(warmup, 3))
    plan.set_timing(True)
    kms = []
    for i in range(3):
        one_step(k)
        kms.append(plan.last_timing()[0])
        k += 1
    plan.set_timing(False)
    kern = ctx.max_over_ranks(statistics.mean(kms))
    ms_per_step = total_ms / steps
    n_cells = nx * ny
    value = n_cells * steps / (total_ms * 1e-3)

    # ---- parity check: records 0..2 from a zero state, owned rows of every rank vs rank 0's 1-rank plan ----
    parity = "pass"
    po = [torch.zeros((4, nb, nx), dtype=torch.float32, device=dev) for _ in range(2)]
    pf = torch.zeros((nb, nx), dtype=torch.uint8, device=dev)
    if rank == 0:
        plan1 = generic.Plan(SBD_VARIANT_PYTHON, nx, ny, nps)
        plan1.set_stream(ctx.stream.cuda_stream)
        plan1.set_static(cdist, dz, ds)
        fo = [torch.zeros((4, ny, nx), dtype=torch.float32, device=dev) for _ in range(2)]
        ff = torch.zeros((ny, nx), dtype=torch.uint8, device=dev)
    for s in range(3):
        th, u, v = recs[s]
        prev = po[(s + 1) % 2]
        plan.step_py(s + 1, cfg.timestep, p, u, v, th, prev[2], prev[3], po[s % 2], flags=pf)
        plan.sync()
        own = torch.cat([po[s % 2][:, r0 - a:r1 - a].reshape(-1), pf[r0 - a:r1 - a].reshape(-1).to(torch.float32)])
        if rank == 0:
            thf, uf, vf = record(0, ny, s)
            prevf = fo[(s + 1) % 2]
            plan1.step_py(s + 1, cfg.timestep, p, uf, vf, thf, prevf[2], prevf[3], fo[s % 2], flags=ff)
            plan1.sync()
            del thf, uf, vf
        if world > 1:
            parts = [torch.empty(5 * (bounds[j + 1] - bounds[j]) * nx, dtype=torch.float32, device=dev)
                     for j in range(world)] if rank == 0 else None
            if rank == 0:
                parts[0] = own
                for j in range(1, world):
                    dist.recv(parts[j], src=j)
            else:
                dist.send(own, dst=0)
        else:
            parts = [own]
        if rank == 0:
            for j in range(world):
                j0, j1 = bounds[j], bounds[j + 1]
                n4 = 4 * (j1 - j0) * nx
                ref = torch.cat([fo[s % 2][:, j0:j1].reshape(-1), ff[j0:j1].reshape(-1).to(torch.float32)])
                # the python variant never processes the global last row (P:165): both sides write the same there
                if not same_bits(parts[j][:n4], ref[:n4]) or not same_bits(parts[j][n4:], ref[n4:]):
                    parity = "FAIL: rows of rank %d differ from the 1-rank plan at record %d" % (j, s)
    if not ctx.all_ok(parity == "pass") and parity == "pass":
        parity = "FAIL on another rank"
    n_trig = int(ctx.sum_over_ranks(int(((pf[r0 - a:r1 - a] & 2) > 0).sum().item())))

    # ---- e2e: one record per call through the host-buffer API ----
    e2e = None
    if not args.no_e2e and not short:
        th, u, v = recs[0]
        hth, hu, hv = (np.array(x.cpu().numpy()) for x in (th, u, v))
        hp = np.array(p.cpu().numpy())
        hws, hwd = np.zeros((nb, nx), np.float32), np.zeros((nb, nx), np.float32)
        hout = np.zeros((4, nb, nx), np.float32)
        for arr in (hth, hu, hv, hout):
            _capi.auto_register(arr)
        for i in range(2):
            plan.step_py(i + 2, cfg.timestep, hp, hu, hv, hth, hws, hwd, hout)
        ctx.barrier()
        ne = 5
        t0 = time.perf_counter()
        for i in range(ne):
            plan.step_py(i + 4, cfg.timestep, hp, hu, hv, hth, hws, hwd, hout)
        torch.cuda.synchronize()
        dt = ctx.max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": n_cells * ne / dt, "unit": UNIT, "steps": ne, "ms_per_step": 1e3 * dt / ne,
               "h2d_bytes_per_step": int(ctx.sum_over_ranks(4 * nb * nx * 5)), "d2h_bytes_per_step": int(ctx.sum_over_ranks(16 * nb * nx)),
               "api": "Plan.step_py per rank with page-locked numpy arrays of its rows (+ overlap): theta, windspeed, winddir "
                      "staged, the selected level of u, v gathered in place, the four result planes copied back"}
    if rank != 0:
        return None
    peak, peak_src = measured_peak()
    abytes = 48.0 * n_cells                       # B_diag(python) = 8 input planes + 4 output planes
    cpu = None
    if not args.no_cpu_baseline and world == 1 and not short:
        cpu = cpu_c4(cfg, land, ice, lon, lat, z, sig)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": max(warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "c4: " + cfg.descr + "; a step is one record, %d records resident and cycled" % nrec,
                   "nx": nx, "ny": ny, "nps": nps, "records": nrec, "get_dist_radius": kk.value, "nn_max": info["nn_max"],
                   "trigger_cells_last_parity_record": n_trig, "rows_per_rank": [bounds[j + 1] - bounds[j] for j in range(world)],
                   "api": "python-variant plan step per record (coast distance / static planes once: sea ice of the synthetic "
                          "records does not change)",
                   "l2": "u, v of one record: %.2f GB per rank" % (2 * 4 * nb * nx * nps / 1e9),
                   "parallelism": ("overlapping latitude row blocks x%d (overlap %d rows >= the largest search radius), "
                                   "no exchange" % (world, overlap)) if world > 1 else "single GPU"},
        "parity_check": parity,
        "roofline": {"bound": "hbm", "kernel": "window kernel of the python variant (k_win4: window sums + wind + thresholds; timed alone, max over ranks)",
                     "achieved": abytes / (kern * 1e-3) / 1e9, "peak": peak * world, "unit": "GB/s",
                     "frac": abytes / (kern * 1e-3) / 1e9 / (peak * world), "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": abytes, "kernel_ms": kern, "kernel_share_of_step": kern / ms_per_step},
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    return line


def cpu_c4(cfg, land, ice, lon, lat, z, sig, rows=300):
    """oracle, literal loops, one record of diag on a band of `rows` rows treated as its own grid"""
    from oracle import oracle as O
    from seabreeze_param_b200 import synth
    a = 300
    sub = synth.Config("c4band", cfg.nx, rows, cfg.nz, 0, "python", cfg.filt, False, cfg.timestep, "band")
    l, i_, zz, ss = (np.ascontiguousarray(x[a:a + rows]) for x in (land, ice, z, sig))
    la = lat[a:a + rows]
    O.set_math(0)
    nthreads = O.max_threads()
    O.set_threads(nthreads)
    coast = O.py_get_edges(l, i_)
    cd = O.py_get_dist(coast, l, lon, la, k=15, literal=False)
    theta, p, u, v = synth.step_fields(sub, 1, l, zz, la)
    tb = []
    for r in range(2):
        t0 = time.perf_counter()
        O.py_diag(1 + r, p, zz, ss, theta, v, u, cd, np.zeros_like(zz), np.zeros_like(zz), timestep=cfg.timestep, literal=True)
        tb.append(time.perf_counter() - t0)
    O.set_math(1)
    O.set_threads(1)
    return {"value": rows * cfg.nx / min(tb), "unit": UNIT, "cores": nthreads, "kind": "port",
            "sample": "one record of diag on a %d-row band (%d x %d cells) of the c4 grid, best of 2, C++ restatement of the "
                      "Fortran (gfortran unavailable), literal growing-window loops, platform libm" % (rows, cfg.nx, rows),
            "ms_per_sample_step": 1e3 * min(tb)}


# ======================================================================================================
# --impl reference
# ======================================================================================================
def run_reference(args, workload):
    """the oracle (C++ restatement, literal mode, platform libm) on the host cores, rank 0 only"""
    from seabreeze_param_b200 import synth
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    from oracle import oracle as O
    cfg = synth.CONFIGS[workload]
    if workload == "c2":
        O.set_math(0)
        nthreads = O.max_threads()
        O.set_threads(nthreads)
        land, ice, lon, lat, z, sig, steps = build_case(cfg, 0)
        h = cfg.h
        coast_h, _ = O.get_edges(ice, land, h)
        cdist = O.get_dist(coast_h, land, lon, lat, 180.0, h, literal=False)
        mk = O.swap_bounds(synth.pad_halo(cdist, h), h)
        zh = O.swap_bounds(synth.pad_halo(z, h), h)
        sh = O.swap_bounds(synth.pad_halo(sig, h), h)
        th = [O.swap_bounds(synth.pad_halo(s[0], h), h) for s in steps]
        st = [np.zeros((cfg.ny, cfg.nx), np.float32) for _ in range(4)]
        n_cells = cfg.nx * cfg.ny
        k = 0
        for _ in range(args.warmup):
            theta, p, u, v = steps[k % NSETS]
            O.seabreeze_diag(cfg.timestep, k + 1, p, u, v, th[k % NSETS], mk, zh, sh, *st, h, literal=True)
            k += 1
        t0 = time.perf_counter()
        for _ in range(args.steps):
            theta, p, u, v = steps[k % NSETS]
            O.seabreeze_diag(cfg.timestep, k + 1, p, u, v, th[k % NSETS], mk, zh, sh, *st, h, literal=True)
            k += 1
        dt = time.perf_counter() - t0
        value = n_cells * args.steps / dt
        sample = ("full c2 grid, %d steps, C++ restatement of the Fortran (gfortran unavailable), platform libm, "
                  "row-parallel std::thread; every call redoes the sigmoid statistics as the reference does" % args.steps)
        ms = 1e3 * dt / args.steps
        descr = "c2: " + cfg.descr
    elif workload == "c3":
        # bounded sample: a 48-row block of the grid's middle, inputs from the host generator at block size
        rows, h, nx = 48, cfg.h, cfg.nx
        sub = synth.Config("c3blk", nx, rows + 2 * h, cfg.nz, h, "generic", cfg.filt, True, cfg.timestep, "c3 block")
        land, ice, lon, lat, z, sig = synth.static_fields(sub)
        O.set_math(0)
        nthreads = O.max_threads()
        O.set_threads(nthreads)
        coast_h, _ = O.get_edges(ice, land, h)
        cdist = O.get_dist(coast_h, land, lon, lat, 180.0, h, literal=False)
        mk = O.swap_bounds(synth.pad_halo(cdist, h), h)[h:-h]       # rows of the padded sub-grid = a halo'd 48-row block
        zh = O.swap_bounds(synth.pad_halo(z, h), h)[h:-h]
        sh = O.swap_bounds(synth.pad_halo(sig, h), h)[h:-h]
        mk, zh, sh = (np.ascontiguousarray(x) for x in (mk, zh, sh))
        theta, p, u, v = synth.step_fields(sub, 1, land, z, lat)
        th = np.ascontiguousarray(O.swap_bounds(synth.pad_halo(theta, h), h)[h:-h])
        p, u, v = (np.ascontiguousarray(x[:, h:h + rows]) for x in (p, u, v))
        st = [np.zeros((rows, nx), np.float32) for _ in range(4)]
        nrun = max(1, min(args.steps, 3))
        for _ in range(1 if args.warmup else 0):
            O.seabreeze_diag(cfg.timestep, 1, p, u, v, th, mk, zh, sh, *st, h, literal=True)
        t0 = time.perf_counter()
        for i in range(nrun):
            O.seabreeze_diag(cfg.timestep, i + 2, p, u, v, th, mk, zh, sh, *st, h, literal=True)
        dt = time.perf_counter() - t0
        value = rows * nx * nrun / dt
        ms = 1e3 * dt / nrun
        sample = ("%d-row block (%d x %d x %d cells, halo %d) of a c3-spaced grid, %d steps (the reference's growing-window "
                  "loops cost tens of seconds per step on the whole grid), C++ restatement of the Fortran (gfortran "
                  "unavailable), platform libm, row-parallel std::thread" % (rows, nx, rows, cfg.nz, h, nrun))
        descr = "c3: " + cfg.descr + " (bounded sample)"
    else:
        land, ice, lon, lat, z, sig = synth.static_fields(cfg)
        c = cpu_c4(cfg, land, ice, lon, lat, z, sig)
        value, ms, sample, nthreads = c["value"], c["ms_per_sample_step"], c["sample"], c["cores"]
        descr = "c4: " + cfg.descr + " (bounded sample)"
    O.set_math(1)
    O.set_threads(1)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak" if workload == "c2" else "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": descr, "nx": cfg.nx, "ny": cfg.ny, "nz": cfg.nz, "halo": cfg.h,
                   "api": "the reference's own routine, full call incl. sigmoid statistics, literal loops"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)   # c2: ~9 ms of device time, the first step's cold start averages out
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=[None, "c2", "c3", "c4"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the short runs of the other north-star configs")
    args = ap.parse_args()
    workload = args.workload or ("c2" if args.gpus == 1 else "c3")
    if args.impl == "reference":
        return run_reference(args, workload)

    ctx = Ctx(args)
    if workload == "c2":
        line = bench_c2(ctx, args)
    elif workload == "c3":
        line = bench_c3(ctx, args, min(args.steps, 40), args.warmup)
    else:
        line = bench_c4(ctx, args, min(args.steps, 96), args.warmup)
    extra = {}
    if args.workload is None and not args.no_extra:
        # the other north-star configs, short: every rank takes part
        ctx.torch.cuda.empty_cache()
        if ctx.world == 1:
            extra["c3"] = bench_c3(ctx, args, 10, 3, short=True)
            ctx.torch.cuda.empty_cache()
        extra["c4"] = bench_c4(ctx, args, 24, 3, short=(ctx.world == 1))
    rc = 0
    if ctx.rank == 0:
        if extra:
            line["other_configs"] = extra
        print(json.dumps(line))
        bad = [line.get("parity_check")] + [e.get("parity_check") for e in extra.values() if e]
        if any(b is not None and b != "pass" for b in bad):
            sys.stderr.write("bench: parity check FAILED: %r\n" % bad)
            rc = 3
    if ctx.world > 1:
        ok = ctx.torch.tensor([rc], device=ctx.dev)
        ctx.dist.broadcast(ok, 0)
        rc = int(ok.item())
        ctx.dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    sys.exit(main())
