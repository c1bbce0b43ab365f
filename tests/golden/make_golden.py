"""Generate the golden input/output vectors under tests/golden/.

Run in the development container, where /root/reference exists:

    make -C oracle ref && python tests/golden/make_golden.py

The expected outputs come from the REAL reference: ``kronmult_omp/kronmult.hpp:77-104`` compiled
unmodified (oracle/_ref/libkronmult_ref_strict.so, single thread so that aliased sums are
deterministic), and -- for the cases small enough -- from the reference's own naive oracle
``tests/utils/kronmult_naive.h:108-121`` (``expected_naive``).  Inputs are seeded N(0,1) like
``tests/utils/data_generation.h:8-19``.  The first two cases are the reference's own correctness
cases ``toy`` and ``small`` (``tests/kronmult_test_gpu.cpp:74-75``: matrix_stride 67, 5 distinct outputs).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from kronmult993_b200 import batch  # noqa: E402
from oracle import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
F64, F32 = torch.float64, torch.float32

# name, d, n, nb, dtype, kwargs
CASES = [
    ("ref_toy_f64", 1, 4, 4, F64, dict(alias="ref", matrices="reftest")),
    ("ref_small_f64", 2, 4, 64, F64, dict(alias="ref", matrices="reftest")),
    ("ref_toy_f32", 1, 4, 4, F32, dict(alias="ref", matrices="reftest")),
    ("ref_small_f32", 2, 4, 64, F32, dict(alias="ref", matrices="reftest")),
    ("ref_medium_f64", 3, 6, 24, F64, dict(alias="ref", matrices="reftest")),
    ("c1_n4d3_f64", 3, 4, 96, F64, dict(alias="distinct")),
    ("c2_n2d2_f64", 2, 2, 257, F64, dict(alias="distinct")),
    ("c3_n4d6_f64", 6, 4, 6, F64, dict(alias="runs", items_per_output=3)),
    ("c4_n8d4_f64", 4, 8, 5, F64, dict(alias="runs", items_per_output=2)),
    ("c5_n4d5_f64", 5, 4, 20, F64, dict(alias="runs", items_per_output=4)),
    ("c5_n4d5_f32", 5, 4, 20, F32, dict(alias="runs", items_per_output=4)),
    ("n4d4_shuffled_f64", 4, 4, 40, F64, dict(alias="shuffled", items_per_output=5)),
    ("n3d3_asgard_f32", 3, 3, 50, F32, dict(alias="shuffled", items_per_output=4, matrices="asgard")),
    ("n5d2_lda9_f64", 2, 5, 33, F64, dict(alias="runs", items_per_output=2, lda=9)),
    ("n7d3_misaligned_f64", 3, 7, 9, F64, dict(alias="ref", nb_distinct=2, lda=11, misalign=1)),
    ("n10d1_f32", 1, 10, 17, F32, dict(alias="ref", nb_distinct=3)),
    ("n2d6_f64", 6, 2, 31, F64, dict(alias="runs", items_per_output=8)),
    ("n9d2_f64", 2, 9, 7, F64, dict(alias="distinct", lda=67)),
    ("n8d3_f32", 3, 8, 6, F32, dict(alias="runs", items_per_output=6)),
]


def main():
    if not oracle.available("ref_strict"):
        raise SystemExit("oracle/_ref/libkronmult_ref_strict.so missing: run `make -C oracle ref` first")
    for i, (name, d, n, nb, dt, kw) in enumerate(CASES):
        p = batch.make_problem(d, n, nb, dt, "cpu", seed=993 + i, **kw)
        hp = p.to_host()
        expected = oracle.run(hp, "ref_strict", threads=1)
        extra = dict(expected=expected)
        if hp.N <= 1024:
            extra["expected_naive"] = oracle.run(hp, "ref_strict", naive=True)
        batch.save_host(os.path.join(HERE, name + ".npz"), hp, **extra)
        print(f"{name}: N={hp.N} nb={nb} {hp.dtype} -> {os.path.getsize(os.path.join(HERE, name + '.npz'))} bytes")


if __name__ == "__main__":
    main()
