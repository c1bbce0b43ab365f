"""Multi-GPU sharding of a kronmult batch by output-pointer ownership (one process per GPU).

No reference counterpart: the reference runs on the current device only
(``kronmult_gpu/kronmult.cu:185,191``).  Batch items only interact through ``+=`` into shared output
vectors (``kronmult.cu:126-129``), so the batch is partitioned such that every output vector has a
single owning rank (``kronmult_partition_by_output`` in ``csrc/partition.cpp``): ranks then run
``kronmult_batched`` on disjoint items *and* disjoint outputs with **no data-path collective**.
Only output groups too large to balance are split across ranks; their per-rank partial sums are
combined with one reduction (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import ctypes
import dataclasses

import numpy as np

from . import api
from .batch import HostProblem


def partition_by_output(out_keys, n_ranks: int, split_threshold: int = 0):
    """``out_keys``: one integer per item identifying its output vector (the output pointer, or any
    group id).  Returns ``(owner int32[nb], needs_reduce uint8[nb])``."""
    lib = api.load_library()
    keys = np.ascontiguousarray(np.asarray(out_keys).astype(np.uint64))
    nb = int(keys.size)
    owner = np.zeros(nb, dtype=np.int32)
    red = np.zeros(nb, dtype=np.uint8)
    code = lib.kronmult_partition_by_output(keys.ctypes.data_as(ctypes.c_void_p), nb, int(n_ranks),
                                            int(split_threshold), owner.ctypes.data_as(ctypes.c_void_p),
                                            red.ctypes.data_as(ctypes.c_void_p))
    if code != 0:
        raise api.KronmultError(code, "kronmult_partition_by_output")
    return owner, red


@dataclasses.dataclass
class Shard:
    problem: HostProblem        # the rank-local problem (local slabs, local offsets)
    items: np.ndarray           # global item indices owned by this rank
    whole_keys: np.ndarray      # global out_off of the output vectors this rank owns outright
    split_keys: np.ndarray      # global out_off of the split output vectors (same on every rank)


def shard_problem(full: HostProblem, rank: int, world: int, split_threshold: int = 0, split_init: str = "zeros"):
    """Rank-local view of ``full``.  Local output slab = [owned outputs with their current values |
    the split outputs].  ``split_init="zeros"``: zero-initialised partial sums (the caller reduces them itself);
    ``"values"``: this rank's copy of every split vector with its current values, as
    ``kronmult_batched_sharded_*`` expects (the library isolates, reduces and adds the partial sums)."""
    owner, red = partition_by_output(full.out_off, world, split_threshold)
    N, d = full.N, full.d
    mine = np.nonzero(owner == rank)[0]
    split_keys = np.unique(full.out_off[red == 1])
    whole_keys = np.unique(full.out_off[mine][red[mine] == 0])
    ar = np.arange(N)
    span = (full.n - 1) * full.lda + full.n
    in_slab = full.in_slab[(full.in_off[mine][:, None] + ar[None, :]).ravel()]
    mo = full.mat_off.reshape(full.nb, d)[mine].ravel()
    mat_slab = full.mat_slab[(mo[:, None] + np.arange(span)[None, :]).ravel()]
    out_whole = full.out_slab[(whole_keys[:, None] + ar[None, :]).ravel()] if whole_keys.size else full.out_slab[:0]
    if split_init == "values" and split_keys.size:
        out_split = full.out_slab[(split_keys[:, None] + ar[None, :]).ravel()]
    else:
        out_split = np.zeros(split_keys.size * N, dtype=full.out_slab.dtype)
    out_slab = np.concatenate([out_whole, out_split])
    lut = {int(k): i * N for i, k in enumerate(whole_keys)}
    lut_s = {int(k): (whole_keys.size + i) * N for i, k in enumerate(split_keys)}
    out_off = np.array([lut_s[int(k)] if red[g] else lut[int(k)] for g, k in zip(mine, full.out_off[mine])],
                       dtype=np.int64)
    local = HostProblem(d, full.n, full.lda, int(mine.size), mat_slab,
                        np.arange(mine.size * d, dtype=np.int64) * span, in_slab,
                        np.arange(mine.size, dtype=np.int64) * N, out_slab, out_off)
    return Shard(local, mine, whole_keys, split_keys), owner, red


def run_shard_on_device(full: HostProblem, rank: int, world: int, comm, device, split_threshold: int = 0, stream=None):
    """One rank of the multi-GPU protocol on a real device: shard, ``kronmult_batched_sharded`` (the library's own
    NCCL all-reduce for the split outputs, owner = rank j % world of the j-th split vector), returns
    ``(shard, local_out_tensor, owner_of_split)``; the owner's copy of a split vector holds the final values."""
    import torch

    from . import batch

    shard, owner, red = shard_problem(full, rank, world, split_threshold, split_init="values")
    p = batch.from_host(shard.problem, device)
    A, i_, o_, w_ = p.pointer_arrays()
    N = full.N
    nw = shard.whole_keys.size
    s = p.out_slab.element_size()
    shared = [p.out_slab.data_ptr() + (nw + j) * N * s for j in range(shard.split_keys.size)]
    split_owner = np.arange(shard.split_keys.size, dtype=np.int32) % world
    api.kronmult_batched_sharded(p.d, p.n, A, p.lda, i_, o_, w_, p.nb, shared, comm, owner=split_owner,
                                 dtype=p.dtype, stream=stream)
    torch.cuda.synchronize(device)
    return shard, p.out_slab, split_owner


def combine_shards(full: HostProblem, shard: Shard, local_out: np.ndarray, dist) -> np.ndarray:
    """Assemble the global output slab from the rank-local results (verification helper).

    The split outputs need the one real collective of the design: a sum-reduction of the per-rank
    partial vectors, added to the original values.  The owned outputs are merely gathered."""
    import torch

    N = full.N
    nw = shard.whole_keys.size
    merged = torch.zeros(full.out_slab.size, dtype=torch.float64)
    ar = np.arange(N)
    if nw:
        idx = (shard.whole_keys[:, None] + ar[None, :]).ravel()
        merged[idx] = torch.from_numpy(np.asarray(local_out[: nw * N], dtype=np.float64))
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(merged)  # gather (disjoint supports)
    if shard.split_keys.size:
        part = torch.from_numpy(np.asarray(local_out[nw * N:], dtype=np.float64).copy())
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(part)  # THE reduction: sum of partial outputs
        idx = (shard.split_keys[:, None] + ar[None, :]).ravel()
        merged[idx] = torch.from_numpy(full.out_slab[idx].astype(np.float64)) + part
    # outputs no item touches keep their values
    touched = np.zeros(full.out_slab.size, dtype=bool)
    touched[(np.unique(full.out_off)[:, None] + ar[None, :]).ravel()] = True
    res = merged.numpy().astype(full.out_slab.dtype)
    res[~touched] = full.out_slab[~touched]
    return res
