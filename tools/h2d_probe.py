"""Concurrent host->device copy bandwidth per GPU (torchrun, one rank per GPU): how much of the e2e number of bench.py
is the box's host fabric.  For k = 1, 2, 4, 8 ... active ranks, every active rank copies a 4 GiB pinned buffer to its GPU
three times at the same moment; rank 0 prints one JSON line per k with the per-GPU and aggregate GB/s (device-event time)."""
import json
import os

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 29  # 4 GiB of float64
h = torch.empty(n, dtype=torch.float64).pin_memory()
h.fill_(1.0)
dv = torch.empty(n, dtype=torch.float64, device="cuda")
k = 1
while k <= world:
    dist.barrier()
    torch.cuda.synchronize()
    ms = 0.0
    if rank < k:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dv.copy_(h, non_blocking=True)  # warm-up
        torch.cuda.synchronize()
    dist.barrier()
    if rank < k:
        e0.record()
        for _ in range(3):
            dv.copy_(h, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        gbs = n * 8 / (t.item() * 1e-3) * 1e-9
        print(json.dumps({"active_gpus": k, "h2d_gbs_per_gpu_slowest": round(gbs, 2), "aggregate_gbs_at_least": round(gbs * k, 1),
                          "buffer_gib_per_gpu": 4, "pinned": True}), flush=True)
    k *= 2
dist.destroy_process_group()
