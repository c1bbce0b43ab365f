"""GPU: the reference's OWN test program, compiled unmodified against include/kronmult.cuh and linked to
libkronmult_b200.so (oracle/Makefile target `reftests`), must pass against this library.

tests/kronmult_test_gpu.cpp:74-75 runs `toy` (n=4,d=1) and `small` (n=4,d=2) with matrix_stride 67 and 5
distinct outputs on managed memory, compares with its naive oracle and prints "Error: <max rel err>";
failure is signalled only by "Test failed!" on stderr when the error exceeds 1e-7 (:51).
The binary is built where /root/reference exists and travels to the GPU box in oracle/_ref/."""
import os
import re
import subprocess

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

BIN = os.path.join(ROOT, "oracle", "_ref", "kronmult_test_gpu")


def test_reference_test_program_passes_against_this_library(kron):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/kronmult_test_gpu not built (needs /root/reference at build time)")
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.dirname(kron.library_path()) + ":" + env.get("LD_LIBRARY_PATH", "")
    res = subprocess.run([BIN], capture_output=True, text=True, timeout=300, env=env)
    assert res.returncode == 0, res.stderr
    assert "Test failed!" not in res.stderr, res.stderr
    errs = [float(x) for x in re.findall(r"^Error: ([0-9.eE+-]+)", res.stdout, flags=re.M)]
    assert len(errs) == 2, res.stdout
    assert all(e <= 1e-7 for e in errs), errs  # the reference's own threshold
    assert all(e <= 1e-13 for e in errs), errs  # and what a correct implementation gives (tests/README.md:19-20)
