"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle.

Tolerances are BASELINE.json's: relative L2 over EVERY output element <= 1e-12 (fp64) / 1e-5 (fp32)
(atomics and reassociation reorder the sums).  Cases follow the reference's own tests
(tests/kronmult_test_gpu.cpp:74-75 toy/small, matrix_stride 67, 5 distinct outputs) and its sweep
envelope (tests/kronmult_fullbench_gpu.cpp:70-74, n in [2,10], d in [1,6]).
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, TOL, golden_files
from kronmult993_b200 import batch

pytestmark = pytest.mark.gpu


def _tol(hp):
    return TOL[str(np.dtype(hp.dtype))]


def _check(kron, oracle_mod, hp, path="auto", expected=None, stream=None):
    p = batch.from_host(hp, "cuda")
    kron.run_problem(p, path=path, stream=stream)
    torch.cuda.synchronize()
    got = p.out_slab.cpu().numpy()
    exp = expected if expected is not None else oracle_mod.run(hp, "oracle", threads=1)
    err = oracle_mod.rel_l2(got, exp)
    assert np.isfinite(got).all()
    assert err <= _tol(hp), f"rel-L2 {err:.3e} > {_tol(hp):.0e} (path {kron.last_path()})"
    return err


@pytest.mark.parametrize("path", ["auto", "generic"])
@pytest.mark.parametrize("fname", golden_files())
def test_golden_vectors(kron, oracle_mod, fname, path):
    hp, extra = batch.load_host(os.path.join(GOLDEN, fname))
    _check(kron, oracle_mod, hp, path, expected=extra["expected"])


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("name", ["toy", "small", "medium"])
def test_reference_named_cases(kron, oracle_mod, name, dt):
    """tests/kronmult_bench_gpu.cpp:68-70 at their full batch sizes (4 / 64 / 384 items)."""
    hp = batch.reference_case(name, dt, "cpu", seed=7).to_host()
    _check(kron, oracle_mod, hp)


def test_reference_large_case_shape(kron, oracle_mod):
    """n = 8, d = 6 (N = 262144 > shared memory): the multi-pass route, 5 distinct outputs."""
    hp = batch.reference_case("large", torch.float64, "cpu", seed=8, nb_cap=12).to_host()
    _check(kron, oracle_mod, hp)
    assert kron.last_path() == "dmma-l2"  # both passes in one persistent kernel, intermediate in L2 (kernel_dmma_l2.cuh)
    kron.set_tuning(12, 0)
    try:
        _check(kron, oracle_mod, hp)
        assert kron.last_path() == "dmma-multipass"  # the two-kernel route through `input`
    finally:
        kron.set_tuning(12, 2)
    _check(kron, oracle_mod, hp, "generic")  # the shape-agnostic multi-pass route must agree too
    assert kron.last_path() == "generic-multipass"


@pytest.mark.parametrize("n,d,nb", [(8, 5, 40), (8, 6, 7), (7, 6, 5), (10, 5, 6), (9, 6, 3), (10, 6, 3), (6, 6, 20)])
def test_vectors_larger_than_shared_memory(kron, oracle_mod, n, d, nb):
    """The top of the reference's sweep envelope (tests/kronmult_fullbench_gpu.cpp:70-74): n^d up to 10^6."""
    for alias, kw in (("runs", dict(items_per_output=3)), ("ref", dict(nb_distinct=2))):
        hp = batch.make_problem(d, n, nb, torch.float64, "cpu", seed=n + d, alias=alias, lda=n + 1, **kw).to_host()
        _check(kron, oracle_mod, hp)


SWEEP = [(n, d) for n in range(2, 11) for d in range(1, 7) if n ** d <= 20000]


@pytest.mark.parametrize("n,d", SWEEP)
def test_sweep_envelope_fp64(kron, oracle_mod, n, d):
    nb = max(3, min(300, 200000 // n ** d))
    hp = batch.make_problem(d, n, nb, torch.float64, "cpu", seed=n * 10 + d, alias="ref", nb_distinct=5,
                            matrices="reftest").to_host()
    _check(kron, oracle_mod, hp)


@pytest.mark.parametrize("n,d", SWEEP)
def test_sweep_envelope_fp32(kron, oracle_mod, n, d):
    """Every shape of the envelope through the AUTOMATIC dispatcher in single precision too (the fp32-only tiny
    shapes (7,2) (9,2) (10,2) (5,3) (3,4) included)."""
    nb = max(3, min(300, 200000 // n ** d))
    hp = batch.make_problem(d, n, nb, torch.float32, "cpu", seed=n * 10 + d, alias="runs", items_per_output=3,
                            lda=n + 3).to_host()
    _check(kron, oracle_mod, hp)


@pytest.mark.parametrize("path,n,d", [("tiny", 2, 2), ("tiny", 4, 2), ("tiny", 3, 2), ("tiny", 9, 1),
                                      ("regtile", 4, 4), ("regtile", 4, 5), ("regtile", 4, 6), ("wspec", 4, 5), ("wspec", 4, 6), ("wspec5", 4, 5), ("sym5", 4, 5), ("sym4", 4, 4),
                                      ("generic", 4, 5), ("generic", 2, 2), ("generic", 8, 4)])
@pytest.mark.parametrize("alias,kw", [("distinct", {}), ("runs", dict(items_per_output=32)),
                                      ("shuffled", dict(items_per_output=5)), ("ref", dict(nb_distinct=1))])
def test_every_kernel_family_every_aliasing(kron, oracle_mod, path, n, d, alias, kw):
    """All output pointers equal (`ref`, 1 distinct) is the worst contention case."""
    N = n ** d
    nb = max(70, min(3000, 600000 // N))
    hp = batch.make_problem(d, n, nb, torch.float64, "cpu", seed=3, alias=alias, **kw).to_host()
    _check(kron, oracle_mod, hp, path)
    assert kron.last_path().startswith(path)


@pytest.mark.parametrize("path", ["sym5", "wspec5"])
@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("nb", [1, 2, 3, 5, 63, 64, 65, 129, 1000, 60000])
def test_n4d5_ragged_batches(kron, oracle_mod, nb, dt, path):
    """n = 4, d = 5: several item streams per CTA -- odd counts, a single item, runs that straddle streams."""
    for alias, kw in (("runs", dict(items_per_output=7)), ("distinct", {})):
        hp = batch.make_problem(5, 4, nb, dt, "cpu", seed=nb, alias=alias, lda=6, **kw).to_host()
        _check(kron, oracle_mod, hp, path)
        assert kron.last_path() == path


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("nb", [1, 2, 3, 31, 32, 33, 64, 127, 129, 300, 4097, 70000])
@pytest.mark.parametrize("d", [4, 5])
def test_n4_half_warp_streams(kron, oracle_mod, nb, dt, d):
    """n = 4, d = 4 (and d = 5 in single precision): two item streams per warp (half-warps) -- odd counts, one item, an
    empty second stream, runs that straddle the streams, and every factor-staging route."""
    if d == 5 and (dt == torch.float64 or nb > 5000):
        pytest.skip("the half-warp kernel covers d = 5 in single precision only")
    for alias, kw, mat in (("runs", dict(items_per_output=7), dict(lda=6)), ("distinct", {}, dict()),
                           ("runs", dict(items_per_output=32), dict(matrices="asgard")),
                           ("ref", dict(nb_distinct=3), dict(matrices="reftest")), ("runs", dict(items_per_output=5), dict(lda=8)),
                           ("runs", dict(items_per_output=4), dict(misalign=1))):
        hp = batch.make_problem(d, 4, nb, dt, "cpu", seed=nb, alias=alias, **kw, **mat).to_host()
        _check(kron, oracle_mod, hp, "sym4")  # forcing the path reaches the d = 5 variant whatever knob 10 says
        assert kron.last_path() == "sym4"


@pytest.mark.parametrize("path", ["sym5", "wspec5"])
@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("layout", ["contiguous", "per_factor", "per_column", "asgard", "lda67", "misaligned"])
def test_n4d5_factor_layouts(kron, oracle_mod, layout, dt, path):
    """Every staging route of the five 4x4 factors: one TMA copy per item (dense, contiguous), one per factor
    (lda = 4, scattered), one per column (16-byte aligned columns), element-wise (anything else)."""
    kw = dict(alias="runs", items_per_output=6)
    if layout == "contiguous":
        hp = batch.make_problem(5, 4, 333, dt, "cpu", seed=5, **kw).to_host()
    elif layout == "per_factor":
        hp = batch.make_problem(5, 4, 333, dt, "cpu", seed=6, **kw).to_host()
        hp.mat_off = np.random.default_rng(0).permutation(hp.mat_off)
    elif layout == "per_column":
        hp = batch.make_problem(5, 4, 333, dt, "cpu", seed=7, lda=8, **kw).to_host()
    elif layout == "asgard":
        hp = batch.make_problem(5, 4, 333, dt, "cpu", seed=8, matrices="asgard", **kw).to_host()
    elif layout == "lda67":
        hp = batch.make_problem(5, 4, 333, dt, "cpu", seed=9, matrices="reftest", **kw).to_host()
    else:
        hp = batch.make_problem(5, 4, 333, dt, "cpu", seed=10, misalign=1, **kw).to_host()
    _check(kron, oracle_mod, hp, path)
    assert kron.last_path() == path


@pytest.mark.parametrize("n,d", [(2, 2), (4, 5), (4, 6), (5, 3), (8, 4)])
def test_unaligned_vectors_and_strided_matrices(kron, oracle_mod, n, d):
    """The API only guarantees alignment to T: shift every slab by one element, lda = 67."""
    for dt in (torch.float64, torch.float32):
        hp = batch.make_problem(d, n, 37, dt, "cpu", seed=11, alias="runs", items_per_output=4, lda=67,
                                misalign=1).to_host()
        _check(kron, oracle_mod, hp)


def test_partially_overlapping_outputs(kron, oracle_mod):
    """The reference is correct for ANY overlap because every add is atomic (kronmult.cu:126-129)."""
    for (n, d) in [(2, 2), (4, 4), (4, 5), (3, 3)]:
        N = n ** d
        p = batch.make_problem(d, n, 40, torch.float64, "cpu", seed=13, alias="distinct")
        hp = p.to_host()
        hp.out_off = (np.arange(40, dtype=np.int64) * (N // 2)) % (hp.out_slab.size - N)  # half-vector steps
        _check(kron, oracle_mod, hp)


def test_edge_batches(kron, oracle_mod):
    hp = batch.make_problem(3, 4, 1, torch.float64, "cpu", seed=1).to_host()
    _check(kron, oracle_mod, hp)
    # nb = 0 is a no-op that returns success (kronmult.cu:191-196 launches an empty grid)
    p = batch.from_host(hp, "cuda")
    before = p.out_slab.clone()
    A, i, o, w = p.pointer_arrays()
    kron.kronmult_batched(3, 4, A, p.lda, i, o, w, 0)
    assert torch.equal(before, p.out_slab)
    # workspace may be NULL here
    kron.kronmult_batched(3, 4, A, p.lda, i, o, None, 1)
    # a d = 0 "product" is the identity on a length-1 vector
    hp0 = batch.make_problem(0, 3, 9, torch.float64, "cpu", seed=2, alias="runs", items_per_output=3).to_host()
    exp = hp0.out_slab.copy()
    np.add.at(exp, hp0.out_off, hp0.in_slab[hp0.in_off])
    _check(kron, oracle_mod, hp0, expected=exp)


def test_invalid_arguments_return_cuda_errors(kron):
    p = batch.make_problem(2, 2, 4, torch.float64, "cuda", seed=1)
    A, i, o, w = p.pointer_arrays()
    with pytest.raises(kron.KronmultError):
        kron.kronmult_batched(2, 2, A, 1, i, o, w, 4)  # lda < n
    with pytest.raises(kron.KronmultError):
        kron.kronmult_batched(40, 2, A, 2, i, o, w, 4)  # 2^40 overflows the reference's int size_input
    with pytest.raises(TypeError):
        kron.kronmult_batched(2, 2, A, 2, i, o, w, 4, dtype=torch.float16)  # only float/double exist


def test_stream_ordered_entry(kron, oracle_mod):
    hp = batch.make_problem(5, 4, 300, torch.float64, "cpu", seed=21, alias="runs", items_per_output=8).to_host()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        _check(kron, oracle_mod, hp, stream=s)


def test_large_batch_tiny_items(kron, oracle_mod):
    """1 Mi items of n = 2, d = 2 (1/16 of BASELINE config 2) against the oracle, every element."""
    hp = batch.make_problem(2, 2, 1 << 20, torch.float64, "cpu", seed=31).to_host()
    _check(kron, oracle_mod, hp)
    assert kron.last_path() == "tiny"


def test_repeated_calls_accumulate(kron, oracle_mod):
    """output += ...: two calls add the product twice (the resident paths never clobber the input)."""
    hp = batch.make_problem(6, 4, 70, torch.float64, "cpu", seed=41, alias="runs", items_per_output=32).to_host()
    p = batch.from_host(hp, "cuda")
    out0 = p.out_slab.clone()
    kron.run_problem(p)
    once = p.out_slab.clone()
    kron.run_problem(p)
    twice = p.out_slab
    d1, d2 = (once - out0), (twice - out0)
    assert float(torch.linalg.norm(d2 - 2 * d1) / torch.linalg.norm(d1)) < 1e-13


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("n,d,nb,r", [(4, 5, 3000, 30), (4, 6, 300, 10), (8, 4, 400, 8), (6, 3, 5000, 25), (4, 4, 6000, 16)])
def test_explicit_plan_on_shuffled_batches(kron, oracle_mod, n, d, nb, r, dt):
    """kronmult_plan_*: a batch whose equal output pointers are scattered is sorted by output pointer; the
    result is the same sum (within tolerance), the number of runs drops to the number of outputs."""
    hp = batch.make_problem(d, n, nb, dt, "cpu", seed=n + d, alias="shuffled", items_per_output=r, lda=n + 2).to_host()
    p = batch.from_host(hp, "cuda")
    A, i, o, w = p.pointer_arrays()
    plan = kron.Plan(d, n, A, p.lda, i, o, nb, dtype=dt)
    st = plan.stats()
    assert st["permuted"] and st["runs_after"] == (nb + r - 1) // r and st["runs_before"] > 2 * st["runs_after"]
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    plan.execute(s)
    s.synchronize()
    exp = oracle_mod.run(hp, "oracle", threads=1)
    assert oracle_mod.rel_l2(p.out_slab.cpu().numpy(), exp) <= _tol(hp)
    # the data may change between executions of the same plan: run again on new inputs (outputs keep accumulating)
    p.in_slab.mul_(-2.0)
    s.wait_stream(torch.cuda.current_stream())  # the update above runs on torch's current stream
    plan.execute(s)
    s.synchronize()
    got2 = p.out_slab.cpu().numpy()
    exp2 = 2 * hp.out_slab - exp  # out0 + P - 2 P
    assert oracle_mod.rel_l2(got2, exp2) <= 10 * _tol(hp)
    plan.destroy()


def test_plan_keeps_grouped_batches_as_they_are(kron, oracle_mod):
    hp = batch.make_problem(5, 4, 2000, torch.float64, "cpu", seed=3, alias="runs", items_per_output=32).to_host()
    p = batch.from_host(hp, "cuda")
    A, i, o, w = p.pointer_arrays()
    plan = kron.Plan(5, 4, A, p.lda, i, o, 2000)
    st = plan.stats()
    assert not st["permuted"] and st["runs_before"] == 63
    plan.execute()
    torch.cuda.synchronize()
    assert oracle_mod.rel_l2(p.out_slab.cpu().numpy(), oracle_mod.run(hp, "oracle", threads=1)) <= 1e-12
    # an empty plan is a no-op
    kron.Plan(5, 4, A, p.lda, i, o, 0).execute()


def test_blocking_entry_plans_shuffled_batches_once(kron, oracle_mod):
    """The drop-in blocking call builds a plan for a scattered batch and reuses it while the pointer arrays stay
    the same; changing one output pointer is detected by the content hash."""
    nb = 8192
    hp = batch.make_problem(5, 4, nb, torch.float64, "cpu", seed=17, alias="shuffled", items_per_output=16).to_host()
    p = batch.from_host(hp, "cuda")
    A, i, o, w = p.pointer_arrays()
    h0, b0 = kron.plan_cache_counters()
    kron.kronmult_batched(5, 4, A, p.lda, i, o, w, nb)
    exp = oracle_mod.run(hp, "oracle", threads=1)
    assert oracle_mod.rel_l2(p.out_slab.cpu().numpy(), exp) <= 1e-12
    h1, b1 = kron.plan_cache_counters()
    assert (h1 - h0, b1 - b0) == (0, 1)
    kron.kronmult_batched(5, 4, A, p.lda, i, o, w, nb)
    h2, b2 = kron.plan_cache_counters()
    assert (h2 - h1, b2 - b1) == (1, 0)
    o[5] = o[6]  # same array, different content -> a new plan, and still the right answer
    hp.out_off[5] = hp.out_off[6]
    before = p.out_slab.cpu().numpy().copy()
    kron.kronmult_batched(5, 4, A, p.lda, i, o, w, nb)
    h3, b3 = kron.plan_cache_counters()
    assert (h3 - h2, b3 - b2) == (0, 1)
    hp.out_slab[:] = before
    exp3 = oracle_mod.run(hp, "oracle", threads=1)
    assert oracle_mod.rel_l2(p.out_slab.cpu().numpy(), exp3) <= 1e-12
    # knob 1 switches implicit planning off
    kron.set_tuning(1, 0)
    try:
        kron.kronmult_batched(5, 4, A, p.lda, i, o, w, nb)
        assert kron.plan_cache_counters() == (h3, b3)
    finally:
        kron.set_tuning(1, 1)


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("n,d", [(n, d) for n in (5, 6, 7, 8) for d in (2, 3)])
def test_dmma_warp_per_item_every_shape(kron, oracle_mod, n, d, dt):
    """kernel_dmma.cuh kron_dmma8s_kernel on every (T, n, d) it is built for (knob 11 = 2; by default only the shapes
    where it measured faster take it): zero-padded 8 x 8 tiles for n < 8, fp32 computed in double, a single item,
    ragged warps, runs that straddle warps, strided / windowed factors, vectors that are not 16-byte aligned."""
    kron.set_tuning(11, 2)
    kron.set_tuning(17, 0)  # the d = 2 lane-per-fibre kernel (rows2) comes first in the automatic dispatch for some of these shapes
    try:
        for alias, kw, extra in (("runs", dict(items_per_output=5), dict(lda=n + 3)),
                                 ("distinct", {}, dict(matrices="reftest")),
                                 ("shuffled", dict(items_per_output=4), dict(misalign=1)),
                                 ("ref", dict(nb_distinct=1), dict(matrices="asgard"))):
            for nb in (1, 3, 130, 1501):
                hp = batch.make_problem(d, n, nb, dt, "cpu", seed=n * 100 + d * 10 + nb % 7, alias=alias, **kw, **extra).to_host()
                _check(kron, oracle_mod, hp)
                assert kron.last_path() == "dmma"
    finally:
        kron.set_tuning(11, 1)
        kron.set_tuning(17, 1)


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("n", [5, 6, 7, 8, 9, 10])
def test_rows2_every_shape(kron, oracle_mod, n, dt):
    """kernel_rows2.cuh kron_rows2_kernel (d = 2, a lane per fibre, floor(32 / n) item slots per warp) on every (T, n) it
    is built for (knob 17 = 2; by default only the shapes where it measured faster take it): a single item, fewer items
    than slots, ragged slots, runs of equal outputs that straddle slots and warps, compact aligned data (16-byte chunk
    copies) as well as strided / windowed factors and vectors that are not 16-byte aligned (element copies)."""
    kron.set_tuning(17, 2)
    try:
        for alias, kw, extra in (("runs", dict(items_per_output=5), {}),
                                 ("runs", dict(items_per_output=7), dict(lda=n + 3)),
                                 ("distinct", {}, dict(matrices="reftest")),
                                 ("shuffled", dict(items_per_output=4), dict(misalign=1)),
                                 ("ref", dict(nb_distinct=1), dict(matrices="asgard"))):
            for nb in (1, 2, 7, 130, 1501, 40007):
                hp = batch.make_problem(2, n, nb, dt, "cpu", seed=n * 100 + nb % 7, alias=alias, **kw, **extra).to_host()
                _check(kron, oracle_mod, hp)
                assert kron.last_path() == "rows2"
    finally:
        kron.set_tuning(17, 1)


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("d,nb", [(6, 1), (6, 2), (6, 4), (6, 5), (6, 9), (6, 13), (6, 24), (6, 41),
                                  (5, 1), (5, 7), (5, 8), (5, 9), (5, 40), (5, 131), (5, 300)])
def test_dmma_l2_persistent_kernel(kron, oracle_mod, d, nb, mode):
    """n = 8, d = 6 (the reference's `large` / `realistic` shape, tests/kronmult_bench_gpu.cpp:71-72) and d = 5 on the
    persistent kernel of kernel_dmma_l2.cuh: chunks of 4 (8) items, a partial last chunk, fewer chunks than ring slots and
    more (the ring wraps at 24 / 128 items), runs of equal outputs that straddle chunks, one output for everything, distinct outputs,
    strided factors, vectors that are not 16-byte aligned; with and without L2 eviction hints (knob 12).  The input
    vectors must come back untouched (this route only reads them)."""
    kron.set_tuning(12, mode)
    try:
        for alias, kw, extra in (("runs", dict(items_per_output=3), dict(lda=11)),
                                 ("ref", dict(nb_distinct=1), {}),
                                 ("distinct", {}, dict(misalign=1))):
            hp = batch.make_problem(d, 8, nb, torch.float64, "cpu", seed=60 + nb, alias=alias, **kw, **extra).to_host()
            p = batch.from_host(hp, "cuda")
            kron.run_problem(p)
            torch.cuda.synchronize()
            assert kron.last_path() == "dmma-l2"
            err = oracle_mod.rel_l2(p.out_slab.cpu().numpy(), oracle_mod.run(hp, "oracle", threads=1))
            assert err <= 1e-12, (alias, nb, err)
            assert np.array_equal(p.in_slab.cpu().numpy(), hp.in_slab)
    finally:
        kron.set_tuning(12, 2)


def test_dmma_l2_launches_on_two_streams_share_the_ring(kron, oracle_mod):
    """Two stream-ordered calls in flight at once: the second waits for the first (event), results stay exact."""
    hps = [batch.make_problem(6, 8, 9, torch.float64, "cpu", seed=70 + i, alias="runs", items_per_output=4).to_host()
           for i in range(2)]
    ps = [batch.from_host(hp, "cuda") for hp in hps]
    arrs = [p.pointer_arrays() for p in ps]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    for rep in range(3):
        for p, (A, i, o, w), st in zip(ps, arrs, streams):
            kron.kronmult_batched(6, 8, A, p.lda, i, o, w, p.nb, stream=st)
    torch.cuda.synchronize()
    for p, hp in zip(ps, hps):
        exp = oracle_mod.run(hp, "oracle", threads=1)
        three = hp.out_slab + 3.0 * (exp - hp.out_slab)
        assert oracle_mod.rel_l2(p.out_slab.cpu().numpy(), three) <= 1e-12


def _pairtile_shapes(dt):
    """(n, d) the pairtile family accepts: the vector (+ stage, + accumulator) fits in shared memory (fp64 tiles of
    n = 9, 10 are shared by two threads)."""
    s = 8 if dt == torch.float64 else 4
    out = []
    for n in range(2, 11):
        for d in range(2, 7):
            if n ** d * s <= 140 * 1024:
                out.append((n, d))
    return out


@pytest.mark.parametrize("n,d", _pairtile_shapes(torch.float64))
def test_pairtile_fp64(kron, oracle_mod, n, d):
    """Every compile-time (n, d) of the pairtile family, forced: ragged streams, runs of equal outputs that straddle
    streams and CTAs, the reference's 5-output pattern with lda = 67, distinct outputs."""
    N = n ** d
    nb = max(7, min(1500, 400000 // N))
    for alias, kw in (("runs", dict(items_per_output=5, lda=n + 1)), ("ref", dict(nb_distinct=5, matrices="reftest")),
                      ("distinct", {})):
        hp = batch.make_problem(d, n, nb, torch.float64, "cpu", seed=100 + n * 7 + d, alias=alias, **kw).to_host()
        _check(kron, oracle_mod, hp, "pairtile")
        assert kron.last_path() == "pairtile"


@pytest.mark.parametrize("n,d", _pairtile_shapes(torch.float32))
def test_pairtile_fp32(kron, oracle_mod, n, d):
    N = n ** d
    nb = max(7, min(1500, 400000 // N))
    for alias, kw in (("runs", dict(items_per_output=5, lda=n + 1)), ("distinct", dict(matrices="reftest"))):
        hp = batch.make_problem(d, n, nb, torch.float32, "cpu", seed=200 + n * 7 + d, alias=alias, **kw).to_host()
        _check(kron, oracle_mod, hp, "pairtile")
        assert kron.last_path() == "pairtile"


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("n,d,nb", [(4, 3, 1), (4, 3, 2), (4, 3, 149), (6, 4, 3), (6, 4, 1000), (5, 2, 1), (5, 2, 4097),
                                    (3, 6, 297), (7, 3, 300), (2, 6, 5000)])
def test_pairtile_ragged_and_misaligned(kron, oracle_mod, n, d, nb, dt):
    """Batch sizes around the stream / CTA boundaries; vectors shifted by one element (element-wise cp.async route)."""
    for mis in (0, 1):
        hp = batch.make_problem(d, n, nb, dt, "cpu", seed=nb + mis, alias="runs", items_per_output=3, lda=n + 3,
                                misalign=mis).to_host()
        _check(kron, oracle_mod, hp, "pairtile")


def _share_inputs(hp, n_unique):
    """Items k and k' with k % n_unique == k' % n_unique read the SAME input vector (what an ASGarD-style caller
    wants: out_i += K_ij x_j for many i per x_j).  Returns the shared problem and an expanded twin with one private
    copy per item, which is what the reference's clobbering contract needs and what the oracle is run on."""
    import copy
    N = hp.N
    shared = copy.copy(hp)
    shared.in_off = hp.in_off[np.arange(hp.nb) % n_unique].copy()
    twin = copy.copy(hp)
    twin.in_slab = np.concatenate([hp.in_slab[o:o + N] for o in shared.in_off])
    twin.in_off = np.arange(hp.nb, dtype=np.int64) * N
    return shared, twin


@pytest.mark.parametrize("n,d,nb", [(2, 2, 500), (4, 3, 300), (3, 5, 200), (4, 4, 200), (4, 5, 333), (4, 6, 40),
                                    (6, 4, 60), (8, 4, 50), (5, 6, 10), (9, 4, 20), (10, 5, 6), (8, 5, 12), (8, 6, 5),
                                    (7, 6, 4)])
def test_read_only_shared_inputs(kron, oracle_mod, n, d, nb):
    """kronmult_batched_const_*: inputs shared between items and left untouched, for every kernel family, including
    the multi-pass routes that otherwise work in place in `input` (they get per-item scratch vectors)."""
    for dt in (torch.float64, torch.float32):
        hp = batch.make_problem(d, n, nb, dt, "cpu", seed=n * d, alias="runs", items_per_output=3, lda=n + 1).to_host()
        shared, twin = _share_inputs(hp, max(1, nb // 4))
        exp = oracle_mod.run(twin, "oracle", threads=1)
        p = batch.from_host(shared, "cuda")
        before = p.in_slab.clone()
        A, i, o, _ = p.pointer_arrays()
        w = None
        if kron.needs_workspace(d, n, dt):
            p.alloc_workspaces()
            A, i, o, w = p.pointer_arrays()
        kron.kronmult_batched_const(d, n, A, p.lda, i, o, w, nb, dtype=dt)
        torch.cuda.synchronize()
        err = oracle_mod.rel_l2(p.out_slab.cpu().numpy(), exp)
        assert err <= _tol(hp), f"rel-L2 {err:.3e} (path {kron.last_path()})"
        assert torch.equal(before, p.in_slab), f"input was modified (path {kron.last_path()})"


def test_read_only_input_needs_workspace_for_multipass(kron):
    """Without scratch vectors a shape that has to go through global memory is refused, not silently clobbered."""
    assert kron.needs_workspace(6, 7, torch.float64) and not kron.needs_workspace(5, 4, torch.float64)
    assert not kron.needs_workspace(6, 8, torch.float64)  # the persistent n = 8, d = 6 kernel only reads `input`
    p = batch.make_problem(6, 7, 3, torch.float64, "cuda", seed=1)
    A, i, o, _ = p.pointer_arrays()
    with pytest.raises(kron.KronmultError):
        kron.kronmult_batched_const(6, 7, A, p.lda, i, o, None, 3, dtype=torch.float64)


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("n,d,nb", [(6, 6, 9), (7, 6, 5), (8, 5, 20), (8, 6, 4), (9, 5, 7), (9, 6, 3), (10, 5, 6),
                                    (10, 6, 3)])
def test_pairtile_multipass(kron, oracle_mod, n, d, nb, dt):
    """Vectors beyond shared memory through the pairtile pass kernels (2 or 3 passes over global memory), forced:
    runs of equal outputs, the reference's few-outputs pattern with lda = 67, ragged unit ranges, misaligned vectors."""
    kron.set_tuning(12, 0)  # (the query below then answers for the multi-kernel routes: n = 8 in double is a pairtile shape too)
    try:
        if not kron.needs_workspace(d, n, dt):
            pytest.skip("fits in shared memory: resident kernel")
        for alias, kw in (("runs", dict(items_per_output=3, lda=n + 1)), ("ref", dict(nb_distinct=2, matrices="reftest")),
                          ("distinct", dict(misalign=1))):
            hp = batch.make_problem(d, n, nb, dt, "cpu", seed=n * 11 + d, alias=alias, **kw).to_host()
            _check(kron, oracle_mod, hp, "pairtile")
            assert kron.last_path() == "pairtile-multipass"
    finally:
        kron.set_tuning(12, 2)


def _fuzz_cases(count, seed):
    rng = np.random.default_rng(seed)
    cases = []
    while len(cases) < count:
        n = int(rng.integers(2, 11))
        d = int(rng.integers(1, 7))
        N = n ** d
        if N > 120000:
            continue
        nb = int(rng.integers(1, max(2, min(700, 300000 // N))))
        alias = str(rng.choice(["runs", "distinct", "shuffled", "ref"]))
        cases.append((n, d, nb, alias, int(rng.integers(1, 9)), int(rng.integers(0, 2)), int(rng.integers(0, 4)),
                      bool(rng.integers(0, 2))))
    return cases


@pytest.mark.parametrize("n,d,nb,alias,r,mis,lda_extra,f32", _fuzz_cases(70, 20261017))
def test_random_shapes_through_the_dispatcher(kron, oracle_mod, n, d, nb, alias, r, mis, lda_extra, f32):
    """Seeded random (n, d, batch size, aliasing pattern, run length, alignment, leading dimension, type) through the
    automatic dispatch, and the same problem through the read-only-input entry point with the inputs left untouched."""
    dt = torch.float32 if f32 else torch.float64
    kw = dict(items_per_output=r) if alias in ("runs", "shuffled") else (dict(nb_distinct=min(5, nb)) if alias == "ref" else {})
    hp = batch.make_problem(d, n, nb, dt, "cpu", seed=n * 100 + d * 10 + nb, alias=alias, lda=n + lda_extra, misalign=mis,
                            **kw).to_host()
    exp = oracle_mod.run(hp, "oracle", threads=1)
    _check(kron, oracle_mod, hp, expected=exp)
    p = batch.from_host(hp, "cuda")
    before = p.in_slab.clone()
    A, i, o, w = p.pointer_arrays()
    if kron.needs_workspace(d, n, dt):
        p.alloc_workspaces()
        A, i, o, w = p.pointer_arrays()
    else:
        w = None
    kron.kronmult_batched_const(d, n, A, p.lda, i, o, w, nb, dtype=dt)
    torch.cuda.synchronize()
    assert oracle_mod.rel_l2(p.out_slab.cpu().numpy(), exp) <= _tol(hp), kron.last_path()
    assert torch.equal(before, p.in_slab)
