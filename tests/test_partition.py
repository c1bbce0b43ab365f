"""CPU-only: the multi-GPU partitioner and the N>1 sharded path (gloo, world_size 2).

Each rank takes the items the partitioner assigns to it, applies the operator to its shard with the
CPU oracle (standing in for the GPU kernel, which is what the ranks run on the B200 box) and the
shards are combined; the result must equal the single-process answer.  Output groups that had to be
split are summed across ranks with one collective, exactly as bench.py does with NCCL.
"""
import ctypes
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from kronmult993_b200 import batch, partition


def test_partition_keeps_groups_together_and_balances():
    g, n_out = batch.output_groups(8192, "runs", items_per_output=32)
    owner, red = partition.partition_by_output(g.numpy(), 8)
    assert red.sum() == 0
    for grp in range(n_out):
        assert len(set(owner[grp * 32:(grp + 1) * 32])) == 1
    counts = np.bincount(owner, minlength=8)
    assert counts.max() - counts.min() <= 32
    assert np.all(np.diff(owner) >= 0)  # fine-grained groups -> contiguous blocks per rank


def test_partition_lpt_for_skewed_groups():
    sizes = [500, 300, 200, 100, 100, 50, 30, 20]
    keys = np.concatenate([np.full(s, i) for i, s in enumerate(sizes)])
    owner, red = partition.partition_by_output(keys, 3)
    loads = np.bincount(owner, minlength=3)
    assert red.sum() == 0 and loads.max() <= 500
    for i in range(len(sizes)):
        assert len(set(owner[keys == i])) == 1


def test_partition_splits_oversized_groups():
    # the reference harness pattern: 5 distinct outputs (tests/kronmult_bench_gpu.cpp:15) on 8 ranks
    g, _ = batch.output_groups(896, "ref", nb_distinct=5)
    owner, red = partition.partition_by_output(g.numpy(), 8, split_threshold=896 // 8)
    counts = np.bincount(owner, minlength=8)
    assert red.sum() > 0 and counts.min() > 0
    assert counts.max() <= 2 * counts.min() + 8


def test_partition_shuffled_and_edge_cases():
    gen = torch.Generator().manual_seed(1)
    g, n_out = batch.output_groups(1000, "shuffled", items_per_output=10, gen=gen)
    owner, _ = partition.partition_by_output(g.numpy(), 4)
    for grp in range(n_out):
        assert len(set(owner[(g == grp).numpy()])) == 1
    owner, _ = partition.partition_by_output(np.zeros(0, dtype=np.int64), 4)
    assert owner.size == 0
    owner, _ = partition.partition_by_output(np.arange(3), 8)  # fewer items than ranks
    assert sorted(owner.tolist()) == sorted(set(owner.tolist()))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, split, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle

    alias = dict(alias="ref", nb_distinct=3) if split else dict(alias="runs", items_per_output=8)
    p = batch.make_problem(3, 4, 96, torch.float64, "cpu", seed=5, **alias)
    full = p.to_host()
    shard, owner, needs_reduce = partition.shard_problem(full, rank, world,
                                                         split_threshold=(96 // world) if split else 0)
    out = oracle.run(shard.problem, "oracle", threads=1) if shard.problem.nb > 0 else shard.problem.out_slab.copy()
    merged = partition.combine_shards(full, shard, out, dist)
    if rank == 0:
        expected = oracle.run(full, "oracle", threads=1)
        q.put(float(oracle.rel_l2(merged, expected)))
    dist.destroy_process_group()


@pytest.mark.parametrize("split", [False, True])
def test_two_rank_sharded_run_matches_single_process(split):
    from oracle import oracle

    oracle.build(("oracle",))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, split, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(120)
        assert pr.exitcode == 0
    assert q.get(timeout=5) < 1e-14
