"""Probe: torch symmetric memory (peer-mapped buffers over NVLink) between 2 ranks."""
import os, sys, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty((4, 1024), dtype=torch.float32, device=dev)
    t.fill_(float(rank + 1))
    hdl = symm.rendezvous(t, dist.group.WORLD)
    torch.cuda.synchronize(); dist.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (4, 1024), torch.float32)
    x = peer.clone(); torch.cuda.synchronize()
    print("rank", rank, "peer value", float(x[0, 0]), "peer ptr", hex(peer.data_ptr()), "own ptr", hex(t.data_ptr()))
    hdl.barrier(channel=0)
    # timing of barrier
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(100): hdl.barrier(channel=0)
    torch.cuda.synchronize()
    print("rank", rank, "symm barrier: %.1f us" % ((time.perf_counter() - t0) / 100 * 1e6))
except Exception as e:
    import traceback; traceback.print_exc()
    print("SYMM FAILED", repr(e))
dist.barrier(); dist.destroy_process_group()
