"""Host logic of the multi-GPU row decomposition used by bench.py (c3 / c4 at N > 1): blocks of equal window work,
boundaries on tile rows, every rank at least one halo tall.  Pure numpy; the device side of the same runs is
verified bit for bit inside every `bench.py --gpus N` run (`parity_check`)."""
import numpy as np
import pytest

import bench


def radius_plane(ny, nx, seed):
    """a coast-like search-radius plane: zero away from a wavy 'coast', growing with distance up to a cap"""
    rng = np.random.default_rng(seed)
    y = np.arange(ny)[:, None]
    x = np.arange(nx)[None, :]
    coast = ny * (0.5 + 0.3 * np.sin(2 * np.pi * x / nx * rng.integers(1, 4)) * rng.uniform(0.3, 1.0))
    d = np.abs(y - coast)
    nn = np.where(d < 72, np.minimum(d + 1, 54), 0).astype(np.int32)
    return nn


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("seed", [1, 2])
def test_balanced_rows_partition_properties(world, seed):
    ny, nx, h = 3000, 400, 54
    nn = radius_plane(ny, nx, seed)
    cost = bench.window_cost_rows(nn, per_cell=60.0)
    b = bench.balanced_rows(cost, world, align=32, min_rows=h)
    assert b[0] == 0 and b[-1] == ny and len(b) == world + 1
    heights = np.diff(b)
    assert (heights >= h).all(), heights                      # a block is at least one halo tall (peer rows exist)
    assert all(r % 32 == 0 for r in b[1:-1]), b                # interior boundaries on tile rows
    work = np.array([cost[b[k]:b[k + 1]].sum() for k in range(world)])
    # the heaviest block is within one tile row of work of the ideal share (boundaries move in steps of 32 rows)
    assert work.max() <= cost.sum() / world + 32 * cost.max() + 1e-6


def test_balanced_rows_uniform_cost_gives_equal_blocks():
    b = bench.balanced_rows(np.ones(1024), 4, align=32, min_rows=6)
    assert b == [0, 256, 512, 768, 1024]


def test_window_cost_counts_window_elements():
    nn = np.array([[0, 1, 2], [3, 0, 0]], np.int32)
    c = bench.window_cost_rows(nn, per_cell=10.0)
    assert c.tolist() == [9 + 25 + 30.0, 49 + 30.0]
