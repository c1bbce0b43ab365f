"""GPU, >= 2 devices: the sharded path on real GPUs with NCCL (skipped on a single-GPU box).

Rank r takes the items the partitioner assigns to it, runs the CUDA library on its shard on cuda:r, and the split
output groups (the reference harness' 5-outputs pattern on more ranks than outputs, tests/kronmult_bench_gpu.cpp:15)
are summed with ONE NCCL all_reduce over device buffers -- the only collective of the design.
The assembled result must equal the single-process CPU oracle.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, split, q):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from kronmult993_b200 import api, batch, partition
    from oracle import oracle

    nb = 480
    alias = dict(alias="ref", nb_distinct=3) if split else dict(alias="runs", items_per_output=8)
    full = batch.make_problem(5, 4, nb, torch.float64, "cpu", seed=5, **alias).to_host()
    shard, owner, red = partition.shard_problem(full, rank, world, split_threshold=(nb // (2 * world)) if split else 0)
    N = full.N
    nw = shard.whole_keys.size
    if shard.problem.nb > 0:
        p = batch.from_host(shard.problem, f"cuda:{rank}")
        api.run_problem(p)
        local = p.out_slab
    else:
        local = torch.from_numpy(shard.problem.out_slab).cuda(rank)
    # assemble on the device: owned outputs have disjoint supports (a sum gathers them); the split outputs are
    # partial sums -> the one real reduction
    merged = torch.zeros(full.out_slab.size, dtype=torch.float64, device=f"cuda:{rank}")
    ar = torch.arange(N, device=merged.device)
    if nw:
        idx = (torch.from_numpy(shard.whole_keys).to(merged.device)[:, None] + ar[None, :]).ravel()
        merged[idx] = local[: nw * N]
    dist.all_reduce(merged)
    if shard.split_keys.size:
        part = local[nw * N:].clone()
        dist.all_reduce(part)
        idx = (torch.from_numpy(shard.split_keys).to(merged.device)[:, None] + ar[None, :]).ravel()
        merged[idx] = torch.from_numpy(full.out_slab).to(merged.device)[idx] + part
    torch.cuda.synchronize()
    if rank == 0:
        expected = oracle.run(full, "oracle", threads=1)
        res = merged.cpu().numpy()
        touched = np.zeros(full.out_slab.size, dtype=bool)
        touched[(np.unique(full.out_off)[:, None] + np.arange(N)[None, :]).ravel()] = True
        res[~touched] = full.out_slab[~touched]
        q.put((float(oracle.rel_l2(res, expected)), int(red.sum()), api.last_path()))
    dist.destroy_process_group()


@pytest.mark.parametrize("split", [False, True])
def test_sharded_run_on_gpus_with_nccl(split):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    from oracle import oracle

    oracle.build(("oracle",))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, split, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(300)
    assert all(pr.exitcode == 0 for pr in procs), [pr.exitcode for pr in procs]
    err, n_split, path = q.get(timeout=10)
    assert err <= 1e-12, err
    assert (n_split > 0) == split
    assert path == "wspec5"
