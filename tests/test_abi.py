"""CPU-only: the C-ABI library loads and exports every symbol the headers declare (no compute)."""
import ctypes
import os
import re
import subprocess

from conftest import ROOT


def _declared(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"//[^\n]*", "", txt)
    return re.findall(r"\b(kronmult_\w+)\s*\(", txt)


def test_library_builds_and_exports_c_abi(kron):
    lib = kron.load_library()
    declared = set(_declared("kronmult_b200.h"))
    assert declared == set(kron.C_SYMBOLS), declared ^ set(kron.C_SYMBOLS)
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in include/kronmult_b200.h but not exported"


def test_cxx_drop_in_symbols_match_reference_mangling(kron):
    """Same mangled names as the reference's libkronmult_gpu (probed with nm)."""
    out = subprocess.run(["nm", "-D", "--defined-only", kron.library_path()], capture_output=True, text=True,
                         check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    for sym in kron.CXX_SYMBOLS:
        assert sym in exported, sym


def test_pow_int_and_version(kron):
    assert kron.pow_int(8, 6) == 262144
    assert kron.pow_int(7, 0) == 1
    assert "sm_100a" in kron.version()


def test_library_is_sm100a_only(kron):
    out = subprocess.run(["cuobjdump", "--list-elf", kron.library_path()], capture_output=True, text=True)
    if out.returncode != 0:
        import pytest
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_reference_signature_header_compiles():
    """include/kronmult.cuh is usable exactly like the reference header (kronmult_gpu/README.md:10,21):
    a consumer TU that calls both specialisations and pow_int compiles and links against the library."""
    import shutil
    import tempfile

    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    src = r'''
#include <kronmult.cuh>
int main() {
    double** pd = nullptr; float** pf = nullptr;
    cudaError e1 = kronmult_batched<double>(1, 2, (double const* const*)pd, 2, pd, pd, pd, 0);
    cudaError e2 = kronmult_batched<float>(1, 2, (float const* const*)pf, 2, pf, pf, pf, 0);
    return (int)e1 + (int)e2 + (pow_int(2, 3) != 8);
}
'''
    from kronmult993_b200 import build
    with tempfile.TemporaryDirectory() as td:
        f = os.path.join(td, "consumer.cu")
        open(f, "w").write(src)
        exe = os.path.join(td, "consumer")
        subprocess.run([nvcc, *build.ARCH, "--std=c++17", "-I", os.path.join(ROOT, "include"), f, "-o", exe,
                        "-L", os.path.dirname(build.LIB), "-lkronmult_b200"], check=True)


def test_needs_workspace_query(kron):
    """kronmult_b200_needs_workspace is host-only: 1 exactly for the shapes whose vector leaves shared memory."""
    import torch

    f64, f32 = torch.float64, torch.float32
    assert not kron.needs_workspace(2, 2, f64)       # tiny
    assert not kron.needs_workspace(5, 4, f64)       # wspec5
    assert not kron.needs_workspace(4, 8, f64)       # dmma
    assert not kron.needs_workspace(5, 6, f64)       # pairtile, 62 KiB
    assert not kron.needs_workspace(6, 6, f32)       # pairtile, 182 KiB single stage
    assert kron.needs_workspace(6, 6, f64)           # pairtile multi-pass
    # dmma-l2: persistent kernel, intermediate in a library-owned ring; the multi-kernel routes (knob 12 = 0) work in place
    assert not kron.needs_workspace(6, 8, f64) and not kron.needs_workspace(5, 8, f64)
    kron.set_tuning(12, 0)
    try:
        assert kron.needs_workspace(6, 8, f64) and kron.needs_workspace(5, 8, f64)
    finally:
        kron.set_tuning(12, 2)
    assert kron.needs_workspace(6, 8, f32)
    assert kron.needs_workspace(5, 10, f32) and kron.needs_workspace(6, 10, f64)
    assert kron.needs_workspace(7, 11, f64)          # outside the envelope: generic multi-pass
    assert not kron.needs_workspace(3, 11, f64)      # generic, resident
    lib = kron.load_library()
    assert lib.kronmult_b200_needs_workspace(3, 4, 2) == -1
