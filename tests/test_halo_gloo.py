"""Multi-rank host logic on CPU (gloo, world_size 2 and 3): the row-block halo exchange reproduces
slices of the single-domain halo fill, and the rank-ordered carry chain of the sigmoid
statistics equals the sequential statistics of the global field."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _chain_host(sigma_h, pass_no, first, carry, h, nx, ny):
    """Host stand-in for sbd_stats_chain_f32 (same contract), sequential float32."""
    a = sigma_h[h:h + ny, h:h + nx].numpy().astype(np.float32).ravel()
    c = carry.astype(np.float32).copy()
    if pass_no == 0:
        s = np.float32(c[0])
        for x in a:
            s = np.float32(s + x)
        c[0] = s
        mx, mn = a.max(), a.min()
        c[2] = mx if first else max(c[2], mx)
        c[3] = mn if first else min(c[3], mn)
    else:
        v = np.float32(c[1]); mean = np.float32(c[4])
        for x in a:
            d = np.float32(x - mean); v = np.float32(v + np.float32(d * d))
        c[1] = v
    return c


def _worker(rank, world, port, nx, ny_loc, h, seed, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from seabreeze_param_b200 import halo
    rng = np.random.default_rng(seed)
    glob = rng.random((ny_loc * world, nx)).astype(np.float32) * 100
    mine = glob[rank * ny_loc:(rank + 1) * ny_loc]
    f = torch.zeros((ny_loc + 2 * h, nx + 2 * h), dtype=torch.float32)
    f[h:h + ny_loc, h:h + nx] = torch.from_numpy(mine)
    ex = halo.RowBlockHalo(nx, ny_loc, h, rank, world, None,
                           chain_fn=lambda s, p, fi, c: _chain_host(s, p, fi, c, h, nx, ny_loc))
    ex.swap_bounds(f)
    stats = ex.global_sigma_stats(f)
    q.put((rank, f.numpy().copy(), stats))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_row_block_halo_and_stats_chain(oracle, world):
    nx, ny_loc, h, seed = 40, 12, 4, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, nx, ny_loc, h, seed, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r, f, st = q.get(timeout=120)
        res[r] = (f, st)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(seed)
    glob = rng.random((ny_loc * world, nx)).astype(np.float32) * 100
    ny = ny_loc * world
    full = np.zeros((ny + 2 * h, nx + 2 * h), np.float32)
    full[h:h + ny, h:h + nx] = glob
    oracle.swap_bounds(full, h)
    _, st_ref = oracle.sigmoid(glob)
    for r in range(world):
        f, st = res[r]
        exp = full[r * ny_loc:r * ny_loc + ny_loc + 2 * h]
        assert np.array_equal(f.view(np.uint32), exp.view(np.uint32)), "rank %d halo'd block differs" % r
        assert np.array_equal(st.view(np.uint32), st_ref[:3].view(np.uint32)), "rank %d global stats differ" % r
