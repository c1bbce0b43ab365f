"""CPU-only: pin the oracle (oracle/kronmult_oracle.c) against the reference.

(1) golden vectors produced by the reference's own kronmult_omp build (tests/golden/make_golden.py),
(2) the reference's naive explicit-Kronecker oracle stored next to them,
(3) bit-for-bit against oracle/_ref/libkronmult_ref_strict.so where that build is present,
(4) the reference's case generator (tests/utils/batch_size.h) and pow_int.
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, TOL, golden_files
from kronmult993_b200 import batch


@pytest.mark.parametrize("fname", golden_files())
def test_oracle_matches_golden_bitwise(oracle_mod, fname):
    hp, extra = batch.load_host(os.path.join(GOLDEN, fname))
    got = oracle_mod.run(hp, "oracle", threads=1)
    # the fixture was produced by the reference header compiled without FMA contraction, single thread
    assert np.array_equal(got, extra["expected"]), fname


@pytest.mark.parametrize("fname", golden_files())
def test_oracle_matches_reference_naive(oracle_mod, fname):
    hp, extra = batch.load_host(os.path.join(GOLDEN, fname))
    if "expected_naive" not in extra:
        pytest.skip("explicit Kronecker matrix too large for this case")
    got = oracle_mod.run(hp, "oracle", threads=1)
    tol = 1e-13 if hp.dtype == np.float64 else 2e-6
    assert oracle_mod.rel_l2(got, extra["expected_naive"]) < tol
    # the restated naive product must reproduce the reference's naive product exactly
    mine = oracle_mod.run(hp, "oracle", naive=True)
    assert np.array_equal(mine, extra["expected_naive"])


def test_oracle_multithreaded_within_tolerance(oracle_mod):
    p = batch.make_problem(4, 4, 512, torch.float64, "cpu", seed=5, alias="runs", items_per_output=32)
    hp = p.to_host()
    a = oracle_mod.run(hp, "oracle", threads=1)
    b = oracle_mod.run(hp, "oracle", threads=4)
    assert oracle_mod.rel_l2(b, a) < 1e-14


@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
@pytest.mark.parametrize("d,n,nb,kw", [
    (3, 4, 200, dict(alias="runs", items_per_output=7)),
    (2, 2, 1000, dict(alias="distinct")),
    (4, 3, 30, dict(alias="shuffled", items_per_output=4, matrices="asgard")),
    (2, 8, 40, dict(alias="ref", matrices="reftest")),
    (1, 1, 5, dict(alias="distinct")),
    (5, 2, 64, dict(alias="ref", nb_distinct=5, lda=5, misalign=1)),
])
def test_oracle_bitwise_vs_reference_build(oracle_mod, dt, d, n, nb, kw):
    if not oracle_mod.available("ref_strict"):
        pytest.skip("oracle/_ref not built here (no /root/reference): golden vectors pin the oracle instead")
    hp = batch.make_problem(d, n, nb, dt, "cpu", seed=17, **kw).to_host()
    mine = oracle_mod.run(hp, "oracle", threads=1)
    ref = oracle_mod.run(hp, "ref_strict", threads=1)
    assert np.array_equal(mine, ref)
    fast = oracle_mod.run(hp, "ref", threads=1)  # the -O3 build used as CPU baseline
    assert oracle_mod.rel_l2(fast, ref) < TOL[str(hp.dtype)] * 0.1


def test_input_and_workspace_clobbering_contract(oracle_mod):
    """kronmult.cuh:23 / kronmult.hpp:47-51: after d passes the result sits in `workspace` when d is
    odd and in `input` when d is even (probed) -- checked on the raw pointer entry."""
    import ctypes

    lib = oracle_mod._lib("oracle")
    for d in (1, 2, 3):
        n, N = 2, 2 ** d
        mats = np.stack([np.diag([1.0, 10.0 ** (j + 1)]) for j in range(d)]).astype(np.float64)
        mats = np.ascontiguousarray(np.transpose(mats, (0, 2, 1)))  # col-major storage
        x = np.ones(N)
        out = np.zeros(N)
        ws = np.zeros(N)
        mp = (ctypes.c_void_p * d)(*[mats[j].ctypes.data for j in range(d)])
        ip = (ctypes.c_void_p * 1)(x.ctypes.data)
        op = (ctypes.c_void_p * 1)(out.ctypes.data)
        wp = (ctypes.c_void_p * 1)(ws.ctypes.data)
        lib.oracle_kronmult_batched_f64(d, n, mp, n, ip, op, wp, 1)
        holder = ws if d % 2 == 1 else x
        assert np.array_equal(holder, out)
        if d == 3:  # last factor acts on the fastest index (kronmult_naive.h:60-61)
            assert out.tolist() == [1, 1000, 100, 1e5, 10, 1e4, 1e3, 1e6]


def test_case_generator_matches_reference(oracle_mod):
    expect = {"toy": 4, "small": 64, "medium": 384, "large": 896, "realistic": 3903}
    for name, (deg, dim, lvl) in batch.REFERENCE_CASES.items():
        assert batch.compute_batch_size(deg, dim, lvl) == expect[name]
        assert oracle_mod.compute_batch_size("oracle", deg, dim, lvl) == expect[name]
        if oracle_mod.available("ref"):
            assert oracle_mod.compute_batch_size("ref", deg, dim, lvl) == expect[name]
    for n in range(1, 11):
        for d in range(0, 7):
            assert oracle_mod.pow_int("oracle", n, d) == n ** d == batch.pow_int(n, d)


def test_reference_alias_rule():
    g, D = batch.output_groups(64, "ref", nb_distinct=5)
    assert D == 5 and g[:5].tolist() == [0, 1, 2, 3, 4]
    assert g[5:].tolist() == [(i * 5) // 64 for i in range(5, 64)]
    g, D = batch.output_groups(4, "ref", nb_distinct=5)  # the reference overflows here
    assert D == 4 and g.tolist() == [0, 1, 2, 3]
