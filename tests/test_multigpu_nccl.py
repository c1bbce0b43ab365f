"""GPU, >= 2 devices: the sharded path on real GPUs with the library's own NCCL collective (skipped on one GPU).

Rank r takes the items the partitioner assigns to it and calls kronmult_batched_sharded_* on cuda:r; the split
output groups (the reference harness' 5-outputs pattern on more ranks than outputs, tests/kronmult_bench_gpu.cpp:15)
are summed INSIDE the call with one ncclAllReduce -- the only collective of the design.
The assembled result must equal the single-process CPU oracle.  torch.distributed only carries the ncclUniqueId and
gathers the pieces for the check.
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


def _worker(rank, world, port, split, n, d, dtname, q):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from kronmult993_b200 import api, batch, partition
    from oracle import oracle

    dt = getattr(torch, dtname)
    nb = 480
    alias = dict(alias="ref", nb_distinct=3) if split else dict(alias="runs", items_per_output=8)
    full = batch.make_problem(d, n, nb, dt, "cpu", seed=5, **alias).to_host()
    comm = api.Comm(dist, device=dev)
    shard, local, split_owner = partition.run_shard_on_device(
        full, rank, world, comm, dev, split_threshold=(nb // (2 * world)) if split else 0)
    ms, ncoll = comm.last_collective()
    N = full.N
    nw = shard.whole_keys.size
    # assemble on the device for the check: every output vector is taken from its one owner (disjoint supports,
    # so a sum gathers them)
    merged = torch.zeros(full.out_slab.size, dtype=torch.float64, device=dev)
    ar = torch.arange(N, device=dev)
    if nw:
        idx = (torch.from_numpy(shard.whole_keys).to(dev)[:, None] + ar[None, :]).ravel()
        merged[idx] = local[: nw * N].double()
    for j, key in enumerate(shard.split_keys):
        if split_owner[j] == rank:
            merged[int(key): int(key) + N] = local[(nw + j) * N: (nw + j + 1) * N].double()
    dist.all_reduce(merged)
    torch.cuda.synchronize()
    if rank == 0:
        expected = oracle.run(full, "oracle", threads=1)
        res = merged.cpu().numpy().astype(full.out_slab.dtype)
        touched = np.zeros(full.out_slab.size, dtype=bool)
        touched[(np.unique(full.out_off)[:, None] + np.arange(N)[None, :]).ravel()] = True
        res[~touched] = full.out_slab[~touched]
        q.put((float(oracle.rel_l2(res, expected)), int(shard.split_keys.size), api.last_path(), ms, ncoll))
    comm.destroy()
    dist.destroy_process_group()


@pytest.mark.parametrize("n,d,dtname,path", [(4, 5, "float64", "sym5"), (8, 4, "float64", "dmma"), (6, 3, "float32", "pairtile")])
@pytest.mark.parametrize("split", [False, True])
def test_sharded_run_on_gpus_with_nccl(split, n, d, dtname, path):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    from oracle import oracle

    oracle.build(("oracle",))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, split, n, d, dtname, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(300)
    assert all(pr.exitcode == 0 for pr in procs), [pr.exitcode for pr in procs]
    err, n_split, last, ms, ncoll = q.get(timeout=10)
    assert err <= (1e-12 if dtname == "float64" else 1e-5), err
    assert (n_split > 0) == split
    assert (ncoll > 0) == split  # the collective runs only when an output group had to be split
    assert last.startswith(path)
    print(f"world={world} split={split} {n=} {d=} {dtname}: rel_l2={err:.2e} collective={ms:.3f} ms x{ncoll}")
