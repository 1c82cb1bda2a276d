"""Host <-> device bandwidth of the box with 1 / 2 / 4 / 8 ranks copying at the same time (torchrun, one rank per GPU): the wall the
end-to-end (host buffers) number of bench.py runs against. Each active rank moves the bench's own sizes -- 1.26 GB of pinned height
maps in, 3.77 GB of pinned RGB out -- on two streams concurrently (PCIe is full duplex), like tx_step_host does.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/pcie_probe.py
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
E = 4096
h_in = torch.empty((E, 240, 320), dtype=torch.float32).pin_memory()
h_out = torch.empty((E, 240, 320, 3), dtype=torch.float32).pin_memory()
d_in = torch.empty_like(h_in, device="cuda")
d_out = torch.empty_like(h_out, device="cuda")
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def run(active: bool, reps: int = 3):
    barrier()
    t0 = time.perf_counter()
    if active:
        for _ in range(reps):
            with torch.cuda.stream(s_in):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s_out):
                h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    return (h_in.numel() * 4 * reps / dt / 1e9, h_out.numel() * 4 * reps / dt / 1e9) if active else (0.0, 0.0)


run(True, 1)  # warm-up (page-locks are touched, copy engines spin up)
res = {}
k = 1
while k <= world:
    h2d, d2h = run(rank < k)
    t = torch.tensor([h2d, d2h], device="cuda", dtype=torch.float64)
    if world > 1:
        lst = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(lst, t)
    else:
        lst = [t]
    if rank == 0:
        act = [x.tolist() for x in lst[:k]]
        res[str(k)] = {"h2d_gbs_per_rank": [round(a[0], 1) for a in act], "d2h_gbs_per_rank": [round(a[1], 1) for a in act],
                       "d2h_gbs_total": round(sum(a[1] for a in act), 1), "h2d_gbs_total": round(sum(a[0] for a in act), 1)}
    k *= 2
if rank == 0:
    out = {"what": "concurrent pinned H2D (1.26 GB) + D2H (3.77 GB) per active rank, GB/s", "world": world, "cpus": os.cpu_count(), "concurrency": res}
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
