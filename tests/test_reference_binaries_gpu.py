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


def _run_ref(kron, name, timeout):
    path = os.path.join(ROOT, "oracle", "_ref", name)
    if not os.path.exists(path):
        pytest.skip(f"oracle/_ref/{name} not built (needs /root/reference at build time)")
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.dirname(kron.library_path()) + ":" + env.get("LD_LIBRARY_PATH", "")
    return subprocess.run([path], capture_output=True, text=True, timeout=timeout, env=env)


def test_reference_bench_program_runs_against_this_library(kron):
    """tests/kronmult_bench_gpu.cpp:68-72: toy, small, medium, large, realistic (stride 67, 5 distinct outputs)."""
    res = _run_ref(kron, "kronmult_bench_gpu", 600)
    assert res.returncode == 0, res.stderr[-2000:]
    got = dict(re.findall(r"^(toy|small|medium|large|realistic): (\d+)ms", res.stdout, flags=re.M))
    assert set(got) == {"toy", "small", "medium", "large", "realistic"}, res.stdout[-2000:]


def test_reference_fullbench_program_runs_against_this_library(kron):
    """tests/kronmult_fullbench_gpu.cpp:70-74: the whole sweep envelope n in [2,10] x d in [1,6] x level in [2,9]
    (432 cases, one blocking call each, managed memory allocated per vector by the reference's harness)."""
    if os.environ.get("KRON_RUN_FULLBENCH") != "1":
        pytest.skip("opt-in (KRON_RUN_FULLBENCH=1): 11 minutes on a B200, nearly all of it the reference harness' "
                    "per-vector cudaMallocManaged calls; log of a full run: profiles/reference_fullbench_gpu_against_b200_r02.txt")
    res = _run_ref(kron, "kronmult_fullbench_gpu", 1500)
    assert res.returncode == 0, res.stderr[-2000:]
    rows = re.findall(r"^degree:(\d+) dimension:(\d+) level:(\d+) batch-size:(\d+): (\d+)ms", res.stdout, flags=re.M)
    assert len(rows) == 9 * 6 * 8, len(rows)
