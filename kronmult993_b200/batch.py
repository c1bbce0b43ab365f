"""Seeded, slab-allocated batch generator for ``kronmult_batched`` problems.

This is the harness-side counterpart of the reference's batch containers
(``tests/utils/utils_gpu.h:37-91`` ``DeviceArrayBatch`` and ``:97-169``
``DeviceArrayBatch_withRepetition``) and of its case generator
(``tests/utils/batch_size.h:8-21``), rebuilt for 10^7-item batches:

* one allocation per *kind* of buffer (inputs, outputs, matrices) instead of one
  ``cudaMallocManaged`` per vector -- 16 Mi allocations are infeasible and 256-byte allocation
  granularity would waste sectors for N = 4;
* reproducible ``N(0,1)`` data from a seed (the reference seeds from ``std::random_device``,
  ``tests/utils/utils_gpu.h:55-56``, so its data cannot be reproduced);
* the reference's aliasing rule ``ptr[i] = ptr[(i*D)/nb]`` (``utils_gpu.h:119-123``) with the
  ``D = min(D, nb)`` fix for its toy-case overflow, plus ASGarD-style contiguous runs of ``r`` items
  per output and a shuffled variant.

The generator works on any torch device (CPU for the oracle-only tests, CUDA for the real thing)
and describes a problem by *element offsets into slabs*; pointer arrays for the C ABI are derived
from them (``KronProblem.pointer_arrays``).  Nothing here computes a Kronecker product.
"""
from __future__ import annotations

import dataclasses
from typing import Optional

import numpy as np
import torch

_SIZEOF = {torch.float64: 8, torch.float32: 4}
_NP = {torch.float64: np.float64, torch.float32: np.float32}


def pow_int(base: int, exponent: int) -> int:
    """``n**d`` as the reference computes it (``kronmult_gpu/kronmult.cu:11-15``), without the
    silent ``int`` overflow: raises instead."""
    v = 1
    for _ in range(exponent):
        v *= base
    if v >= 2**31:
        raise OverflowError(f"{base}^{exponent} does not fit the reference's int size_input")
    return v


def compute_batch_size(degree: int, dimension: int, grid_level: int, nb_distinct_outputs: int = 5) -> int:
    """Batch count of a reference test/bench case (``tests/utils/batch_size.h:8-21``)."""
    n, d = degree, dimension
    N = pow_int(n, d)
    cap = (395_000_000_000 - nb_distinct_outputs * N) // (N * (2 + d * n * n))
    formula = (2**grid_level) * (grid_level ** min(1, d - 1))
    return int(min(cap, formula))


# the five named reference cases: tests/kronmult_bench_gpu.cpp:68-72 (degree, dimension, level)
REFERENCE_CASES = {
    "toy": (4, 1, 2),
    "small": (4, 2, 4),
    "medium": (6, 3, 6),
    "large": (8, 6, 7),
    "realistic": (8, 6, 9),
}


@dataclasses.dataclass
class HostProblem:
    """A problem held in numpy arrays (what the CPU oracle consumes)."""

    d: int
    n: int
    lda: int
    nb: int
    mat_slab: np.ndarray
    mat_off: np.ndarray  # int64 [nb*d], element offsets into mat_slab
    in_slab: np.ndarray
    in_off: np.ndarray  # int64 [nb]
    out_slab: np.ndarray
    out_off: np.ndarray  # int64 [nb]

    @property
    def N(self) -> int:
        return pow_int(self.n, self.d)

    @property
    def dtype(self):
        return self.in_slab.dtype


@dataclasses.dataclass
class KronProblem:
    """A problem held in torch tensors on one device."""

    d: int
    n: int
    lda: int
    nb: int
    dtype: torch.dtype
    mat_slab: torch.Tensor
    mat_off: torch.Tensor  # int64 [nb*d]
    in_slab: torch.Tensor
    in_off: torch.Tensor  # int64 [nb]
    out_slab: torch.Tensor
    out_off: torch.Tensor  # int64 [nb]
    n_outputs: int
    unique_matrices: int
    ws_slab: Optional[torch.Tensor] = None  # only the reference CUDA kernel needs real workspaces
    ws_off: Optional[torch.Tensor] = None

    @property
    def N(self) -> int:
        return pow_int(self.n, self.d)

    @property
    def device(self):
        return self.in_slab.device

    # ---- bookkeeping shared by bench.py and DESIGN.md (SURVEY.md §8d definitions) ----
    def flops(self) -> int:
        return self.nb * 2 * self.d * self.n ** (self.d + 1)

    def algorithmic_bytes(self) -> int:
        s = _SIZEOF[self.dtype]
        N = self.N
        return (self.nb * N * s + self.unique_matrices * self.n * self.n * s
                + 2 * self.n_outputs * N * s + self.nb * (self.d + 2) * 8)

    # ---- views ----
    def pointer_arrays(self):
        """(A, in, out, ws) int64 tensors of raw addresses on ``device``: the four pointer arrays
        of ``kronmult_batched`` (``kronmult_gpu/kronmult.cuh:28-32``).  ``ws`` points every item at
        one shared dummy vector unless real workspaces were allocated."""
        s = _SIZEOF[self.dtype]
        A = self.mat_slab.data_ptr() + self.mat_off * s
        i = self.in_slab.data_ptr() + self.in_off * s
        o = self.out_slab.data_ptr() + self.out_off * s
        if self.ws_slab is not None and self.ws_off is not None:
            w = self.ws_slab.data_ptr() + self.ws_off * s
        else:
            if self.ws_slab is None:
                self.ws_slab = torch.zeros(self.N + 2, dtype=self.dtype, device=self.device)
            w = torch.full((self.nb,), self.ws_slab.data_ptr(), dtype=torch.int64, device=self.device)
        return A.contiguous(), i.contiguous(), o.contiguous(), w.contiguous()

    def alloc_workspaces(self):
        self.ws_slab = torch.zeros(self.nb * self.N, dtype=self.dtype, device=self.device)
        self.ws_off = torch.arange(self.nb, dtype=torch.int64, device=self.device) * self.N

    def to_host(self) -> HostProblem:
        c = lambda t: t.detach().cpu().numpy().copy()
        return HostProblem(self.d, self.n, self.lda, self.nb, c(self.mat_slab), c(self.mat_off), c(self.in_slab),
                           c(self.in_off), c(self.out_slab), c(self.out_off))

    def select_outputs_to_host(self, groups: torch.Tensor) -> "tuple[HostProblem, torch.Tensor]":
        """Host copy of exactly the items that feed the output vectors ``groups`` (indices into
        the output slab in units of N) -- the strided-subset parity check of SURVEY.md §8d for
        configs too large for host RAM.  Returns the compacted problem and the selected groups."""
        N = self.N
        groups = groups.to(self.device)
        gid = self.out_off // N
        lut = torch.full((int(self.out_slab.numel() // N) + 1,), -1, dtype=torch.int64, device=self.device)
        lut[groups] = torch.arange(groups.numel(), device=self.device)
        sel = lut[gid]
        items = torch.nonzero(sel >= 0).flatten()
        nb = int(items.numel())
        ar = torch.arange(N, device=self.device)
        in_sub = self.in_slab[(self.in_off[items][:, None] + ar[None, :]).flatten()]
        out_sub = self.out_slab[(groups[:, None] * N + ar[None, :]).flatten()]
        span = (self.n - 1) * self.lda + self.n
        mo = self.mat_off.view(self.nb, self.d)[items].flatten()
        am = torch.arange(span, device=self.device)
        mat_sub = self.mat_slab[(mo[:, None] + am[None, :]).flatten()]
        hp = HostProblem(self.d, self.n, self.lda, nb, mat_sub.cpu().numpy(),
                         (torch.arange(nb * self.d) * span).numpy().astype(np.int64), in_sub.cpu().numpy(),
                         (torch.arange(nb) * N).numpy().astype(np.int64), out_sub.cpu().numpy(),
                         (sel[items] * N).cpu().numpy().astype(np.int64))
        return hp, groups


def _randn_into(t: torch.Tensor, gen: torch.Generator, chunk: int = 1 << 28):
    flat = t.view(-1)
    for s in range(0, flat.numel(), chunk):
        e = min(flat.numel(), s + chunk)
        flat[s:e].normal_(generator=gen)


def output_groups(nb: int, alias: str, items_per_output: int = 1, nb_distinct: int = 5,
                  gen: Optional[torch.Generator] = None, device="cpu") -> "tuple[torch.Tensor, int]":
    """Map item -> output vector index.

    ``distinct``  every item its own output;
    ``runs``      contiguous runs of ``items_per_output`` items share an output (ASGarD-style);
    ``shuffled``  the ``runs`` map under a random permutation of the items (exercises the
                  unsorted path);
    ``ref``       the reference harness rule (``tests/utils/utils_gpu.h:112-123``): items
                  0..D-1 own outputs 0..D-1 and item i >= D uses output (i*D)/nb.
    """
    k = torch.arange(nb, dtype=torch.int64, device=device)
    if alias == "distinct":
        return k, nb
    if alias in ("runs", "shuffled"):
        r = max(1, int(items_per_output))
        g = k // r
        if alias == "shuffled":
            perm = torch.randperm(nb, generator=gen, device=device)
            g = g[perm]
        return g, (nb + r - 1) // r
    if alias == "ref":
        D = min(int(nb_distinct), nb)
        g = (k * D) // nb
        g[:D] = k[:D]
        return g, D
    raise ValueError(f"unknown aliasing map {alias!r}")


def make_problem(d: int, n: int, nb: int, dtype=torch.float64, device="cpu", seed: int = 993, *,
                 alias: str = "distinct", items_per_output: int = 1, nb_distinct: int = 5,
                 matrices: str = "dense", lda: Optional[int] = None, misalign: int = 0,
                 init: bool = True) -> KronProblem:
    """Build one seeded problem.

    ``matrices``: ``dense`` -- each item owns ``d`` compact matrices (``lda = n`` unless given);
    ``reftest`` -- each matrix is its own ``n*lda`` block with ``lda = 67``, the layout of
    ``tests/kronmult_test_gpu.cpp:21,30``; ``asgard`` -- ``n x n`` windows into ``d`` shared
    ``L x L`` coefficient matrices, ``L = 64 n``, ``lda = L`` (many items share matrix data).
    ``misalign`` shifts every slab by that many elements so that vectors are element-aligned but
    not 16-byte aligned (the API only promises alignment to ``T``).
    """
    N = pow_int(n, d)
    device = torch.device(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))

    def slab(numel):
        t = torch.empty(numel + misalign, dtype=dtype, device=device)
        v = t[misalign:]
        if init:
            _randn_into(v, gen)
        return v

    g, n_out = output_groups(nb, alias, items_per_output, nb_distinct, gen, device)
    in_slab = slab(nb * N)
    in_off = torch.arange(nb, dtype=torch.int64, device=device) * N
    out_slab = slab(n_out * N)
    out_off = g * N

    if matrices == "dense":
        lda_ = n if lda is None else int(lda)
        block = n * lda_
        mat_slab = slab(nb * d * block)
        mat_off = torch.arange(nb * d, dtype=torch.int64, device=device) * block
        uniq = nb * d
    elif matrices == "reftest":
        lda_ = 67 if lda is None else int(lda)
        block = n * lda_
        mat_slab = slab(nb * d * block)
        mat_off = torch.arange(nb * d, dtype=torch.int64, device=device) * block
        uniq = nb * d
    elif matrices == "asgard":
        L = 64 * n
        lda_ = L
        mat_slab = slab(d * L * L)
        rb = torch.randint(0, 64, (nb * d,), generator=gen, device=device)
        cb = torch.randint(0, 64, (nb * d,), generator=gen, device=device)
        j = torch.arange(nb * d, dtype=torch.int64, device=device) % d
        mat_off = j * (L * L) + rb * n + cb * (n * L)
        uniq = min(nb * d, d * 64 * 64)
    else:
        raise ValueError(f"unknown matrix mode {matrices!r}")
    if lda_ < n:
        raise ValueError("matrix_stride must be >= matrix_size")

    return KronProblem(d, n, lda_, nb, dtype, mat_slab, mat_off, in_slab, in_off, out_slab, out_off, n_out, uniq)


def reference_case(name: str, dtype=torch.float64, device="cpu", seed: int = 993, nb_distinct: int = 5,
                   nb_cap: Optional[int] = None) -> KronProblem:
    """One of the reference's named cases (``tests/kronmult_bench_gpu.cpp:68-72``) with its
    ``matrix_stride = 67`` (``:21``) and 5 distinct outputs (``:15``)."""
    degree, dimension, level = REFERENCE_CASES[name]
    nb = compute_batch_size(degree, dimension, level, nb_distinct)
    if nb_cap is not None:
        nb = min(nb, nb_cap)
    return make_problem(dimension, degree, nb, dtype, device, seed, alias="ref", nb_distinct=nb_distinct,
                        matrices="reftest")


def from_host(hp: HostProblem, device="cuda") -> KronProblem:
    """Upload a numpy problem (e.g. a golden fixture) to a torch device."""
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    dtype = torch.float64 if hp.in_slab.dtype == np.float64 else torch.float32
    N = hp.N
    n_out = int(hp.out_slab.size // N)
    return KronProblem(hp.d, hp.n, hp.lda, hp.nb, dtype, t(hp.mat_slab), t(hp.mat_off.astype(np.int64)),
                       t(hp.in_slab), t(hp.in_off.astype(np.int64)), t(hp.out_slab), t(hp.out_off.astype(np.int64)),
                       n_out, int(np.unique(hp.mat_off).size))


def save_host(path: str, hp: HostProblem, **extra) -> None:
    np.savez_compressed(path, d=hp.d, n=hp.n, lda=hp.lda, nb=hp.nb, mat_slab=hp.mat_slab, mat_off=hp.mat_off,
                        in_slab=hp.in_slab, in_off=hp.in_off, out_slab=hp.out_slab, out_off=hp.out_off, **extra)


def load_host(path: str) -> "tuple[HostProblem, dict]":
    z = np.load(path)
    hp = HostProblem(int(z["d"]), int(z["n"]), int(z["lda"]), int(z["nb"]), z["mat_slab"], z["mat_off"],
                     z["in_slab"], z["in_off"], z["out_slab"], z["out_off"])
    extra = {k: z[k] for k in z.files if k not in ("d", "n", "lda", "nb", "mat_slab", "mat_off", "in_slab",
                                                    "in_off", "out_slab", "out_off")}
    return hp, extra
