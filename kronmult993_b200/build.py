"""In-tree build of libkronmult_b200.so (hand-written CUDA for sm_100a, no JIT cache).

The shared object is written next to this file so that it travels with the repository snapshot to
the GPU box; it is git-ignored.  ``nvcc`` cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libkronmult_b200.so")
MICROBENCH = os.path.join(_HERE, "kron_microbench")
OBJ = os.path.join(_HERE, "_obj")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "--std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", *ARCH,
              *os.environ.get("KRON_EXTRA_NVCC_FLAGS", "").split()]  # e.g. -DKRON_WSPEC5_EXPERIMENTS (ablation builds)


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (no fallback exists)")


def _newer(target: str, sources) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _sources():
    out = []
    for root in (CSRC, os.path.join(os.path.dirname(_HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h", ".cpp")):
                out.append(os.path.join(root, f))
    return out


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Build (if stale) under an exclusive file lock: torchrun starts one process per GPU at the same moment, and
    eight concurrent rebuilds of the same object files hand a half-linked .so to whoever loads first."""
    import fcntl

    srcs = _sources()
    if not force and _newer(LIB, srcs):
        return LIB
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and _newer(LIB, srcs):  # another process built it while this one waited
                return LIB
            return _build_library_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_library_locked(verbose: bool) -> str:
    units = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))
             if f.endswith((".cu", ".cpp")) and not f.startswith("microbench")]
    # one nvcc process per translation unit, all at once (the pairtile instantiations dominate), then link
    os.makedirs(OBJ, exist_ok=True)
    procs = []
    for u in units:
        o = os.path.join(OBJ, os.path.splitext(os.path.basename(u))[0] + ".o")
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", u, "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((o, cmd, subprocess.Popen(cmd, cwd=CSRC)))
    for o, cmd, pr in procs:
        if pr.wait() != 0:
            raise subprocess.CalledProcessError(pr.returncode, cmd)
    # link to a temporary name and rename: a concurrent loader sees the old library or the new one, never a torso
    tmp = LIB + ".tmp"
    subprocess.run([_nvcc(), *ARCH, "-shared", "-o", tmp, *[o for o, _, _ in procs], "-lpthread", "-ldl"], check=True, cwd=CSRC)
    os.replace(tmp, LIB)
    return LIB


def build_microbench(force: bool = False) -> str:
    src = os.path.join(CSRC, "microbench.cu")
    if not os.path.exists(src):
        return ""
    if not force and _newer(MICROBENCH, [src]):
        return MICROBENCH
    subprocess.run([_nvcc(), *NVCC_FLAGS, "-o", MICROBENCH, src], check=True, cwd=CSRC)
    return MICROBENCH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
    print(build_microbench(force=True))
