"""swap_bounds across GPUs: the reference's halo_exchange_mod::swap_bounds (generic/
halo_exchange_mod.f90:12-17 is an empty stub; the UM's real one is MPI, UM/vn10.7/
sea_breeze_diag.F90:408-410) as NCCL send/recv between latitude row blocks.

Decomposition: 1-D row blocks along latitude, one process per GPU.  A halo'd
block is (ny_loc + 2h, nx + 2h); halo rows are contiguous, so each exchange is two sends and two
receives of h*(nx+2h) floats.  East-west periodic wrap stays on the GPU (one kernel,
sbd_swap_bounds_dev_f32); domain edges (first/last rank) replicate the edge row.

The sigmoid statistics (generic/sea_breeze_diag.f90:466-478) are sequential fp32 accumulations over
the *global* grid in memory order; with row blocks every rank's chunk is contiguous in that
order, so the chain is a rank-ordered pipeline of carries (global_sigma_stats), done once per
plan.
"""
import numpy as np

from . import _capi


class RowBlockHalo:
    def __init__(self, nx, ny_local, h, rank=0, world=1, device=None, group=None, chain_fn=None):
        self.nx, self.ny, self.h = nx, ny_local, h
        self.rank, self.world = rank, world
        self.device = device
        self.group = group
        self.lib = None
        self.chain_fn = chain_fn  # tests inject a host implementation of the carry chain (gloo, CPU)
        self.up = rank - 1 if rank > 0 else None       # owns the rows above (smaller latitude index)
        self.down = rank + 1 if rank < world - 1 else None
        self.ns_mask = (1 if self.up is None else 0) | (2 if self.down is None else 0)
        self.launches_per_step = 1

    # -- halo fill of one halo'd field (torch CUDA tensor, (ny+2h, nx+2h), contiguous float32) --
    def _local_fill(self, field):
        """East-west periodic wrap on every row, edge-row replication on the domain's own north /
        south edges.  CUDA tensors: one kernel (sbd_swap_bounds_dev_f32).  CPU tensors (only the
        gloo tests of the exchange logic use them): the same pure data movement by slicing."""
        h, ny, nx = self.h, self.ny, self.nx
        if field.is_cuda:
            import torch
            if self.lib is None:
                self.lib = _capi.load()
            stream = torch.cuda.current_stream(field.device).cuda_stream
            _capi.check(self.lib.sbd_swap_bounds_dev_f32(field.data_ptr(), nx, ny, h, self.ns_mask, stream),
                        "swap_bounds")
            return
        import torch
        idx = (torch.arange(-h, nx + h) % nx) + h
        field[h:h + ny, :] = field[h:h + ny, :][:, idx]
        if self.ns_mask & 1:
            field[0:h, :] = field[h:h + 1, :]
        if self.ns_mask & 2:
            field[ny + h:, :] = field[ny + h - 1:ny + h, :]

    def halo_mode(self):
        """Bit mask for Plan.set_halo_mode: E-W periodic wrap and the replicated domain-edge rows are
        folded into the kernels' reads; only rows owned by a neighbour rank are exchanged."""
        return 1 | (self.ns_mask << 1)

    def swap_bounds(self, field, local_fill=True):
        """local_fill=False: the plan runs with halo_mode() (virtual halos), so only the NCCL
        exchange of neighbour-owned rows is left to do."""
        h, ny = self.h, self.ny
        if local_fill:
            self._local_fill(field)
        if self.world > 1:
            import torch.distributed as dist
            ops = []
            if self.up is not None:
                ops.append(dist.P2POp(dist.isend, field[h:2 * h], self.up, self.group))
                ops.append(dist.P2POp(dist.irecv, field[0:h], self.up, self.group))
            if self.down is not None:
                ops.append(dist.P2POp(dist.isend, field[ny:ny + h], self.down, self.group))
                ops.append(dist.P2POp(dist.irecv, field[ny + h:ny + 2 * h], self.down, self.group))
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return field

    def swap_bounds_step(self, field):
        """Per-timestep halo refresh of theta for a plan in halo_mode(): exchange only."""
        return self.swap_bounds(field, local_fill=False)

    # -- global sequential statistics of sigma over all ranks' interiors (rank-ordered carry) --
    def global_sigma_stats(self, sigma_h):
        """Returns float32[3] = mean, 2/sqrt(var/n), (max-min)/4 of the global interior, evaluated in
        the reference's sequential order (rank-major == global memory order)."""
        import torch
        import torch.distributed as dist
        h, ny, nx = self.h, self.ny, self.nx
        ldx = nx + 2 * h
        interior = sigma_h.data_ptr() + 4 * (h * ldx + h)
        dev = sigma_h.device
        carry = np.zeros(5, dtype=np.float32)
        n_glob = np.float32(nx * ny * self.world)

        def chain(pass_no):
            nonlocal carry
            buf = torch.zeros(5, dtype=torch.float32, device=dev)
            if self.up is not None:
                dist.recv(buf, self.up, group=self.group)
                carry = buf.cpu().numpy().copy()
            first = 1 if (self.up is None and pass_no == 0) else 0
            if self.chain_fn is not None:
                carry = self.chain_fn(sigma_h, pass_no, first, carry)
            else:
                if self.lib is None:
                    self.lib = _capi.load()
                _capi.check(self.lib.sbd_stats_chain_f32(interior, nx, ny, ldx, pass_no, first, carry.ctypes.data),
                            "stats_chain")
            if self.down is not None:
                dist.send(torch.from_numpy(carry).to(dev), self.down, group=self.group)
            # the last rank holds the global carry: broadcast it
            out = torch.from_numpy(carry).to(dev)
            dist.broadcast(out, self.world - 1, group=self.group)
            return out.cpu().numpy().copy()

        c0 = chain(0)
        mean = np.float32(c0[0] / n_glob)
        carry = c0.copy()
        carry[1] = np.float32(0.0)
        carry[4] = mean
        if self.up is not None:
            pass  # carry is overwritten by the received one in chain(1)
        c1 = chain(1)
        var = np.float32(c1[1])
        std = np.float32(np.float32(2.0) / np.sqrt(np.float32(var / n_glob)))
        r = np.float32(np.float32(c0[2] - c0[3]) / np.float32(4.0))
        return np.array([mean, std, r], dtype=np.float32)


class PeerHalo:
    """Per-step halo rows without a copy: every rank keeps its halo'd theta planes in symmetric
    memory (torch.distributed._symmetric_memory: peer-mapped allocations over NVLink / NVSwitch),
    and the window kernel of a neighbour reads the rows it needs straight from them
    (Plan.set_peer_theta).  What is left per step is a device-side barrier over the symmetric
    memory's signal pads ("my theta for this step is in place / I am done reading yours"), or --
    cheaper -- the post / wait words of sync_words() handled inside the window kernel itself
    (Plan.set_peer_sync): no launch, and only the CTAs that touch neighbour rows ever wait."""

    def __init__(self, nbuf, ny_local, nx, h, rank, world, device, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.rank, self.world, self.ny, self.h = rank, world, ny_local, h
        shape = (nbuf, ny_local + 2 * h, nx + 2 * h)
        self.buf = symm.empty(shape, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.up = self.hdl.get_buffer(rank - 1, shape, torch.float32) if rank > 0 else None
        self.dn = self.hdl.get_buffer(rank + 1, shape, torch.float32) if rank < world - 1 else None
        plane_bytes = 4 * shape[1] * shape[2]
        self._peers = [(self.up.data_ptr() + i * plane_bytes if self.up is not None else None, ny_local,
                        self.dn.data_ptr() + i * plane_bytes if self.dn is not None else None) for i in range(nbuf)]

        # in-kernel post / wait words (Plan.set_peer_sync): [0] written by the northern neighbour,
        # [1] by the southern one, [2], [3] time-out markers
        self.flags = symm.empty((16,), dtype=torch.int32, device=device)
        self.flags.zero_()
        self.fhdl = symm.rendezvous(self.flags, group if group is not None else dist.group.WORLD)
        self.flags_up = self.fhdl.get_buffer(rank - 1, (16,), torch.int32) if rank > 0 else None
        self.flags_dn = self.fhdl.get_buffer(rank + 1, (16,), torch.int32) if rank < world - 1 else None
        torch.cuda.synchronize(device)
        dist.barrier(group)  # everybody's words are zero before anybody posts

    def sync_words(self):
        """(local, northern neighbour's, southern neighbour's) flag arrays for Plan.set_peer_sync"""
        return self.flags, self.flags_up, self.flags_dn

    def timed_out(self):
        """True if a window kernel gave up waiting for a neighbour's post (sbd.h: sync_local[2], [3])."""
        return bool((self.flags[2:4] != 0).any().item())

    def plane(self, i):
        return self.buf[i]

    def peers(self, i):
        """(address of the northern neighbour's plane i or None, its interior rows, address of the
        southern neighbour's plane i or None) -- the arguments of Plan.set_peer_theta"""
        return self._peers[i]

    def barrier(self):
        self.hdl.barrier(channel=0)


class PyRowBlock:
    """Row-block sharding of the PYTHON variant (diag, seabreeze_diag_python.f90:49-285) -- no exchange at all.

    The python variant has no halos: its index map clamps rows at the array edge (P:201) and wraps
    columns (P:202).  A rank that owns rows [r0, r1) of the (nlats, nlons) grid therefore builds its
    plan on rows [a, b) = [max(0, r0 - overlap), min(nlats, r1 + overlap)) of every 2-D field and
    restricts the trigger to its own rows (Plan.set_row_window): with overlap >= the largest search
    radius no owned cell's window ever reaches the block's edge, except where that edge is the
    domain's own edge and clamps exactly as the single-domain run does.  The sigmoid statistics are
    taken from the whole grid (sequential order), the rule "the last row is never processed"
    (P:165) falls on the global last row because only the last block contains it.  Per record each
    rank reads its rows (+ overlap) of theta, u, v from the input files on its own: the path shards
    with overlapping reads and no collective (config c4 of BASELINE.json).
    """

    def __init__(self, nlons, nlats, nps, rank, world, cdist, z, std, overlap, params=None, device=None):
        """cdist, z, std: the GLOBAL (nlats, nlons) float32 fields (numpy, C order == Fortran (nlons, nlats));
        overlap: rows read beyond the owned block, >= the largest search radius (<= k + 1, k the
        get_dist radius; `required_overlap` checks it after the fact)."""
        from . import generic
        self.nx, self.ny, self.nz = nlons, nlats, nps
        self.rank, self.world = rank, world
        per = (nlats + world - 1) // world
        self.r0, self.r1 = min(rank * per, nlats), min((rank + 1) * per, nlats)
        self.a, self.b = max(0, self.r0 - overlap), min(nlats, self.r1 + overlap)
        self.overlap = overlap
        # global sequential statistics of std (P:291-309) -- static, every rank computes the same bits
        st = np.zeros(4, dtype=np.float32)
        generic.sigmoid(np.ascontiguousarray(std, dtype=np.float32), stats=st)
        self.stats = st[:3].copy()
        self.plan = generic.Plan(generic.SBD_VARIANT_PYTHON, nlons, self.b - self.a, nps, params=params)
        self.plan.set_row_window(self.r0 - self.a, self.r1 - self.a)
        cut = lambda f: np.ascontiguousarray(f[self.a:self.b], dtype=np.float32)
        self.plan.set_static(cut(cdist), cut(z), cut(std), stats=self.stats)

    def required_overlap(self):
        """Largest search radius among the owned cells; the overlap must not be smaller."""
        return self.plan.info()["nn_max"]

    def rows(self):
        """(a, b): the rows of the global 2-D fields this rank has to read per record."""
        return self.a, self.b

    def step(self, timestep_number, timestep_min, p, u_blk, v_blk, theta_blk, ws_blk, wd_blk, flags=None):
        """One record.  u_blk, v_blk: (nps, b-a, nlons); theta_blk, ws_blk, wd_blk: (b-a, nlons) -- rows
        [a, b) of the global arrays.  Returns (sb_con, t0, windspeed, winddir) for the OWNED rows
        [r0, r1) only, each (r1-r0, nlons)."""
        nb = self.b - self.a
        out = np.zeros((4, nb, self.nx), dtype=np.float32)
        self.plan.step_py(timestep_number, timestep_min, p, u_blk, v_blk, theta_blk, ws_blk, wd_blk, out, flags=flags)
        lo, hi = self.r0 - self.a, self.r1 - self.a
        return tuple(out[k, lo:hi] for k in range(4))
