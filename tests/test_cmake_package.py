"""CPU-only (nvcc cross-compiles): the drop-in CMake packaging of SURVEY.md section 8(f) rank 4.

1. configure + build the repository's CMakeLists.txt (target `kronmult_gpu`, the reference's target name,
   kronmult_gpu/CMakeLists.txt:6-11) and, where the reference checkout exists, its three OWN GPU programs
   tests/kronmult_{test,bench,fullbench}_gpu.cpp UNMODIFIED against it (tests/CMakeLists.txt:52-65);
2. `cmake --install` into a scratch prefix;
3. a consumer project does find_package(kronmult_gpu) and links kronmult_gpu::kronmult_gpu through <kronmult.cuh>.
Nothing is executed (no GPU here); tests/test_reference_binaries_gpu.py runs the programs on the B200 box.
"""
import os
import shutil
import subprocess
import tempfile

import pytest

from conftest import ROOT

REF = "/root/reference"

CONSUMER_CMAKE = """cmake_minimum_required(VERSION 3.18)
set(CMAKE_CUDA_ARCHITECTURES 100a)
project(consumer LANGUAGES CXX CUDA)
find_package(kronmult_gpu REQUIRED)
add_executable(consumer consumer.cu)
target_link_libraries(consumer PRIVATE kronmult_gpu::kronmult_gpu)
"""
CONSUMER_CU = r"""
#include <kronmult.cuh>
#include <kronmult_b200.h>
int main() {
    double** pd = nullptr;
    cudaError e = kronmult_batched<double>(1, 2, (double const* const*)pd, 2, pd, pd, pd, 0);
    return (int)e + (pow_int(2, 3) != 8) + (kronmult_pow_int(3, 2) != 9);
}
"""


def _run(cmd, cwd, timeout=1500):
    res = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True, timeout=timeout)
    assert res.returncode == 0, f"{' '.join(cmd)}\n{res.stdout[-3000:]}\n{res.stderr[-3000:]}"
    return res.stdout


def test_cmake_build_install_and_find_package():
    cmake = shutil.which("cmake")
    if not cmake or not (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        pytest.skip("cmake / nvcc unavailable")
    gen = ["-G", "Ninja"] if shutil.which("ninja") else []
    env_nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    with tempfile.TemporaryDirectory() as td:
        bld, pre, con = (os.path.join(td, x) for x in ("build", "prefix", "consumer"))
        args = [cmake, "-S", ROOT, "-B", bld, *gen, "-DCMAKE_BUILD_TYPE=Release", f"-DCMAKE_CUDA_COMPILER={env_nvcc}",
                f"-DCMAKE_INSTALL_PREFIX={pre}", "-DCMAKE_CXX_COMPILER=/usr/bin/g++", "-DCMAKE_CUDA_HOST_COMPILER=/usr/bin/g++"]
        have_ref = os.path.exists(os.path.join(REF, "tests", "kronmult_test_gpu.cpp"))
        if have_ref:
            args.append(f"-DKRONMULT_REFERENCE_DIR={REF}")
        _run(args, td)
        _run([cmake, "--build", bld, "-j", str(os.cpu_count() or 4)], td)
        assert os.path.exists(os.path.join(bld, "libkronmult_b200.so"))
        if have_ref:
            for prog in ("kronmult_test_gpu", "kronmult_bench_gpu", "kronmult_fullbench_gpu"):
                assert os.path.exists(os.path.join(bld, prog)), prog
        _run([cmake, "--install", bld], td)
        assert os.path.exists(os.path.join(pre, "include", "kronmult.cuh"))
        assert os.path.exists(os.path.join(pre, "lib", "cmake", "kronmult_gpu", "kronmult_gpuConfig.cmake"))
        os.makedirs(con)
        open(os.path.join(con, "CMakeLists.txt"), "w").write(CONSUMER_CMAKE)
        open(os.path.join(con, "consumer.cu"), "w").write(CONSUMER_CU)
        _run([cmake, "-S", con, "-B", os.path.join(con, "b"), *gen, f"-DCMAKE_PREFIX_PATH={pre}",
              f"-DCMAKE_CUDA_COMPILER={env_nvcc}", "-DCMAKE_CXX_COMPILER=/usr/bin/g++",
              "-DCMAKE_CUDA_HOST_COMPILER=/usr/bin/g++"], td)
        _run([cmake, "--build", os.path.join(con, "b")], td)
        out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(pre, "lib", "libkronmult_b200.so")],
                             capture_output=True, text=True).stdout
        assert "_Z16kronmult_batchedIdE9cudaErroriiPKPKT_iPPS1_S7_S7_i" in out and "kronmult_batched_sharded_f64" in out
