#!/usr/bin/env python
"""bench.py -- grid-cells/s of the sea-breeze trigger step (BASELINE.json metric) on N B200s.

A "step" is one call of seabreeze_diag (generic/sea_breeze_diag.f90:55-271) over one timestep of
synthetic fields on the N512-class grid of BASELINE.json configs[1] (1024 x 769 x 70, halo 6).

  value      device-resident: inputs already in HBM, plan API (static fields processed once per
             plan), K steps back to back over rotating input sets (> L2), CUDA
             events on the launching stream around the whole timed region.
  e2e        the same step through the public host-buffer API (numpy arrays in pinned memory):
             H2D of p,u,v,theta,state + kernels + D2H of state/sb_con/flags inside the timed region.
  roofline   the dominant kernel, the level scan k_scan_quad_regs: its algorithmic bytes nC (4 nz + 28)
             (its share of B_diag = 20N + nC(4nz+24)) / its launch time measured with
             CUDA events by the library, against MEASURED_PEAKS.json hbm_gbs; traffic = DRAM bytes per
             launch from the committed ncu launch list (profiles/traffic_k_scan_quad_regs.json).
  cpu_baseline  the oracle (C++ restatement of the Fortran; gfortran does not exist here), literal
             loops, platform libm, all host threads, full c2 grid, best of 3 steps.

N > 1 (torchrun): weak scaling -- every rank owns one N512-sized latitude block of a taller grid; the
halo rows of theta are read from the neighbour ranks' planes over NVLink (symmetric memory) with the
per-step handshake inside the window kernel (BENCH_PEER_SYNC=barrier: a barrier launch per step;
BENCH_HALO=nccl: NCCL send/recv beside the level scan).  See seabreeze_param_b200/halo.py.

--impl reference: times the oracle (kind "port") on the host cores for the same config.
"""
import argparse
import json
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
NSETS = int(os.environ.get("BENCH_NSETS", "4"))  # rotating input sets: 4 x 661 MB of p,u,v >> 126 MB L2


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


def run_reference(args, cfg):
    """--impl reference: the oracle (C++ restatement, literal mode) on the host cores."""
    from oracle import oracle as O
    from seabreeze_param_b200 import synth
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    O.set_math(0)  # platform libm: what the gfortran build of the reference calls
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
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "c2: " + cfg.descr, "nx": cfg.nx, "ny": cfg.ny, "nz": cfg.nz, "halo": cfg.h,
                   "api": "seabreeze_diag (generic), full call incl. sigmoid statistics, literal loops"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port",
                         "sample": "full c2 grid, %d steps, C++ restatement of the Fortran (gfortran "
                                   "unavailable), platform libm, row-parallel std::thread" % args.steps},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)   # ~9 ms of device time: the first step's cold start averages out
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    from seabreeze_param_b200 import synth
    cfg = synth.CONFIGS[args.workload]
    if cfg.variant != "generic":
        raise SystemExit("bench.py times the generic seabreeze_diag step; pick a generic workload")
    if args.impl == "reference":
        return run_reference(args, cfg)

    import torch
    import torch.distributed as dist
    from seabreeze_param_b200 import generic, halo
    from seabreeze_param_b200._capi import SBD_VARIANT_GENERIC, load

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = load()
    assert lib.sbd_device_count() >= 1

    # ---------------- inputs (host) -> HBM ----------------
    h, nx, ny, nz = cfg.h, cfg.nx, cfg.ny, cfg.nz
    land, ice, lon, lat, z, sig, steps = build_case(cfg, rank)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    # static chain on the GPU: get_edges -> get_dist -> halo'd planes (dummy_model.f90:32-36)
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
    # With several ranks theta lives in symmetric (peer-mapped) memory: the window kernel reads the
    # halo rows owned by a neighbour rank straight from that rank's plane over NVLink.
    # BENCH_HALO=nccl selects the send/recv exchange instead.
    use_peer = world > 1 and os.environ.get("BENCH_HALO", "peer") == "peer"
    peer = None
    if use_peer:
        try:
            peer = halo.PeerHalo(NSETS, ny, nx, h, rank, world, dev)
        except Exception as e:  # no symmetric memory on this box: fall back to NCCL send/recv
            sys.stderr.write("bench: symmetric memory unavailable (%r), using NCCL halo exchange\n" % (e,))
            use_peer = False
        ok = torch.tensor([1 if peer is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            peer, use_peer = None, False
    dsteps = []
    for i, (theta, p, u, v) in enumerate(steps):
        if peer is not None:
            th = peer.plane(i)
            th[h:h + ny, h:h + nx] = t(theta)
        else:
            th = haloed(t(theta))
        dsteps.append((th, t(p), t(u), t(v)))
    if peer is not None:
        torch.cuda.synchronize()
        dist.barrier()
    state = [torch.zeros((ny, nx), dtype=torch.float32, device=dev) for _ in range(4)]
    flags = torch.zeros((ny, nx), dtype=torch.uint8, device=dev)

    plan = generic.Plan(SBD_VARIANT_GENERIC, nx, ny, nz, h)
    # everything below is enqueued on one dedicated stream; torch events are recorded on it too
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    torch.cuda.set_stream(stream)
    plan.set_stream(stream.cuda_stream)
    # swap_bounds of the per-step field theta is folded into the kernels' reads (E-W periodic wrap,
    # domain-edge rows); between ranks the N-S halo rows still travel by NCCL every step
    plan.set_halo_mode(ex.halo_mode())
    t_static0 = time.perf_counter()
    stats = ex.global_sigma_stats(sh) if world > 1 else None
    plan.set_static(mk, zh, sh, stats=stats)
    torch.cuda.synchronize()
    static_ms = 1e3 * (time.perf_counter() - t_static0)
    info = plan.info()
    n_cells = nx * ny
    n_coastal = info["n_coastal"]

    # BENCH_PIPE=1: consecutive steps overlap (window kernel of step k+1 under the level scan of step
    # k; all input sets are resident before the loop starts, which is what sbd_plan_set_pipelined
    # asks for).  Measured on c2: 56.7 us per step against 44.2 us unpipelined -- the scan's CTAs
    # take the SMs first and leave the window kernel half its residency -- so it is off by default.
    pipelined = os.environ.get("BENCH_PIPE", "0") != "0"
    plan.set_pipelined(pipelined)
    # multi-GPU: the per-step "theta in place" handshake runs inside the window kernel (post / wait
    # words in symmetric memory); BENCH_PEER_SYNC=barrier uses a symmetric-memory barrier launch
    peer_barrier = os.environ.get("BENCH_PEER_SYNC", "flags") == "barrier"
    if peer is not None and not peer_barrier:
        plan.set_peer_sync(*peer.sync_words())
    comm = torch.cuda.Stream(device=dev) if world > 1 else None
    ev_prev = [None]

    def one_step(k):
        """One timestep.  theta's halo is a per-step input: the E-W periodic wrap and the domain-edge
        rows are read in place by the kernels (plan halo mode), N-S rows owned by neighbour ranks
        arrive by NCCL send/recv on a separate stream beside the level scan, which needs no halos."""
        th, p, u, v = dsteps[k % NSETS]
        if world == 1:
            plan.step(cfg.timestep, k + 2, p, u, v, th, *state, flags=flags)
            return
        if peer is not None:
            # symmetric-memory path: one device-side barrier ("every rank's theta of this step is in
            # place, nobody still reads the previous one"), then the ordinary two-kernel step whose
            # window kernel reads the neighbours' rows over NVLink -- no copy, no extra stream
            if peer_barrier:
                peer.barrier()
            plan.set_peer_theta(*peer.peers(k % NSETS))
            plan.step(cfg.timestep, k + 2, p, u, v, th, *state, flags=flags)
            return
        # NCCL path: the exchange runs on its own stream beside the level scan
        plan.step_scan(p, u, v)
        if ev_prev[0] is not None:
            comm.wait_event(ev_prev[0])
        with torch.cuda.stream(comm):
            ex.swap_bounds_step(th)  # NCCL send/recv of the N-S halo rows
            ev_h = comm.record_event()
        stream.wait_event(ev_h)
        plan.step_finish(cfg.timestep, k + 2, th, *state, flags=flags)
        ev_prev[0] = stream.record_event()

    k = 0
    for _ in range(max(args.warmup, 3)):
        one_step(k); k += 1
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # ---------------- timed region: K steps, CUDA events on the launching stream ----------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]  # start / end only: nothing sits between the steps
    l0 = lib.sbd_launch_count()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    ev[0].record(stream)
    for i in range(args.steps):
        one_step(k)
        k += 1
    ev[1].record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = lib.sbd_launch_count() - l0
    total_ms = ev[0].elapsed_time(ev[-1])
    # per-kernel durations (roofline leg): same steps again with CUDA events around each kernel,
    # recorded by the library on the launching stream
    plan.set_timing(True)
    win_ms, kern_ms = [], []
    for i in range(min(args.steps, 20)):
        one_step(k)
        a_ms, b_ms, _ = plan.last_timing()
        win_ms.append(a_ms); kern_ms.append(b_ms)
        k += 1
    plan.set_timing(False)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
        cc = torch.tensor([n_coastal], dtype=torch.float64, device=dev)
        dist.all_reduce(cc)
        n_coastal_all = int(cc.item())
    else:
        n_coastal_all = n_coastal
    ms_per_step = total_ms / args.steps
    value = world * n_cells * args.steps / (total_ms * 1e-3)

    # ---------------- e2e: host (pinned) buffers through the public API ----------------
    e2e = None
    if not args.no_e2e:
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
        hsteps = []
        for i in range(2):
            theta, p, u, v = steps[i]
            th_h = synth.pad_halo(theta, h)
            hsteps.append((pin(th_h), pin(p), pin(u), pin(v)))
        hstate = [pin(np.zeros((ny, nx), np.float32)) for _ in range(4)]
        hflags = torch.zeros((ny, nx), dtype=torch.uint8).pin_memory().numpy()
        ne = max(3, min(args.steps, 10))
        for i in range(2):
            th_h, p, u, v = hsteps[i % 2]
            plan.step(cfg.timestep, i + 2, p, u, v, th_h, *hstate, flags=hflags)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for i in range(ne):
            th_h, p, u, v = hsteps[i % 2]
            plan.step(cfg.timestep, i + 4, p, u, v, th_h, *hstate, flags=hflags)  # blocking: outputs are on the host
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        # p,u,v sit in pinned host memory.  p: the northern rows of every level travel by DMA, the
        # southern quads' columns are gathered in place by the level scan (zero-copy, 16 B per quad
        # and level); u,v: one level per quad, zero-copy; theta and the three state planes are
        # staged by cudaMemcpyAsync; the four result planes + flags come back
        nq, nqn, ysp = info["n_quads"], info["n_quads_north"], info["host_split_row"]
        h2d = (4 * ysp * nx * nz + 16 * (nq - nqn) * nz + 16 * nq * 2
               + 4 * ((nx + 2 * h) * (ny + 2 * h) + 3 * n_cells))
        d2h = 4 * (4 * n_cells) + n_cells
        e2e = {"value": world * n_cells * ne / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": ne, "ms_per_step": 1e3 * dt / ne,
               "api": "Plan.step with numpy arrays in pinned host memory (blocking call); p: DMA of the rows that "
                      "hold coastal-zone cells (up to the split row) + zero-copy gather of the columns south of it",
               "host_input_bytes_per_step": 4 * (3 * n_cells * nz + (nx + 2 * h) * (ny + 2 * h) + 3 * n_cells)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak()
    six_frac = sum(1 for i in range(args.steps) if ((k - args.steps + i + 2) * cfg.timestep) % 21600 == 0) / args.steps
    # k_scan_quad's share of B_diag: per coastal cell all nz levels of p, one level of u and v, the
    # previous ws, wd and n_thc in, sb_con and ws (wd on a 6-hourly step) out = nC (4 nz + 28);
    # the window kernel owns the remaining 20 N - 4 nC
    abytes = n_coastal * (4.0 * nz + 28.0 + 4.0 * six_frac)
    abytes_step = algorithmic_bytes(n_cells, n_coastal, nz, six_frac)
    kms = statistics.mean(kern_ms)
    achieved = abytes / (kms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "sbdk::k_scan_quad_regs<5,128> (level scan over p + thresholds; timed alone.  The window kernel "
                "sbdk::k_trigger, timed alone: %.2f us; in the timed steps the two overlap)" % (1e3 * statistics.mean(win_ms)), "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": abytes, "algorithmic_bytes_per_step": abytes_step,
                "step_achieved_gbs": abytes_step / (ms_per_step * 1e-3) / 1e9,
                "step_frac": abytes_step / (ms_per_step * 1e-3) / 1e9 / peak, "kernel_ms": kms,
                "kernel_ms_min": min(kern_ms), "kernel_share_of_step": kms / ms_per_step}
    prof = os.path.join(ROOT, "profiles", "traffic_k_scan_quad_regs.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            pass

    cpu = None
    if not args.no_cpu_baseline:
        from oracle import oracle as O
        O.set_math(0)
        nthreads = O.max_threads()
        O.set_threads(nthreads)
        mk_h, zh_h, sh_h = (x.cpu().numpy() for x in (mk, zh, sh))
        st = [np.zeros((ny, nx), np.float32) for _ in range(4)]
        reps = 3
        tb = []
        for i in range(reps):
            theta, p, u, v = steps[i % NSETS]
            th_h = dsteps[i % NSETS][0].cpu().numpy()
            t0 = time.perf_counter()
            O.seabreeze_diag(cfg.timestep, i + 1, p, u, v, th_h, mk_h, zh_h, sh_h, *st, h, literal=True)
            tb.append(time.perf_counter() - t0)
        cpu = {"value": n_cells / min(tb), "unit": UNIT, "cores": nthreads, "kind": "port",
               "sample": "full c2 grid, best of %d steps, C++ restatement of the Fortran (gfortran unavailable), "
                         "literal loops, platform libm" % reps, "ms_per_step": 1e3 * min(tb)}
        O.set_math(1)
        O.set_threads(1)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "c2: " + cfg.descr + (" per GPU (latitude blocks, %d ranks)" % world if world > 1 else ""),
                   "nx": nx, "ny": ny * world, "nz": nz, "halo": h, "n_coastal_cells": n_coastal_all,
                   "api": "plan step%s (static fields processed once per plan: %.1f ms, SURVEY 8d)" % (
                       ", pipelined (window kernel of step k+1 runs under the level scan of step k)" if pipelined else "", static_ms),
                   "l2": "rotating %d input sets of p,u,v (%.0f MB each) > 126 MB L2; no flush kernel in the timed region"
                         % (NSETS, 3 * 4 * n_cells * nz / 1e6),
                   "parallelism": ("row-block x%d, theta halo rows %s" % (
                       world, "read from the neighbours' planes over NVLink (symmetric memory), " + ("per-step device barrier" if peer_barrier else "post/wait flags inside the window kernel")
                       if use_peer else "via NCCL send/recv")) if world > 1 else "single GPU"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "static_setup_ms": static_ms,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
