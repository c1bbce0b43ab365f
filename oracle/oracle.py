"""ctypes front-end of the CPU checker -- TEST INFRASTRUCTURE ONLY.

Loads ``oracle/liboracle.so`` (the C restatement, ``oracle/kronmult_oracle.c``) and, when present,
``oracle/_ref/libkronmult_ref*.so`` (the UNMODIFIED reference ``kronmult_omp/kronmult.hpp:77-104``
compiled by ``oracle/Makefile``).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import this module; the product package
``kronmult993_b200`` never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {
    "oracle": os.path.join(_HERE, "liboracle.so"),
    "ref": os.path.join(_HERE, "_ref", "libkronmult_ref.so"),
    "ref_strict": os.path.join(_HERE, "_ref", "libkronmult_ref_strict.so"),
}
_PREFIX = {"oracle": "oracle", "ref": "ref", "ref_strict": "ref"}
_loaded: dict = {}

_LL = ctypes.POINTER(ctypes.c_longlong)


def build(targets=("oracle", "ref", "refgpu")) -> None:
    """Run the committed recipe.  ``ref``/``refgpu`` are no-ops where /root/reference is absent."""
    subprocess.run(["make", "-s", "-C", _HERE, *targets], check=True)


def available(which: str) -> bool:
    return os.path.exists(_LIBS[which])


def _lib(which: str):
    if which not in _loaded:
        path = _LIBS[which]
        if not os.path.exists(path):
            if which == "oracle":
                build(("oracle",))
            else:
                raise FileNotFoundError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        _loaded[which] = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
    return _loaded[which]


def _sfx(dtype) -> str:
    dt = np.dtype(dtype)
    if dt == np.float64:
        return "f64"
    if dt == np.float32:
        return "f32"
    raise TypeError(f"kronmult is defined for float32/float64 only, got {dt}")


def _p(a: np.ndarray, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def pow_int(which: str, a: int, b: int) -> int:
    f = getattr(_lib(which), f"{_PREFIX[which]}_pow_int")
    f.restype = ctypes.c_int
    return f(ctypes.c_int(a), ctypes.c_int(b))


def compute_batch_size(which: str, degree: int, dimension: int, level: int, nb_distinct: int = 5) -> int:
    f = getattr(_lib(which), f"{_PREFIX[which]}_compute_batch_size")
    f.restype = ctypes.c_int
    return f(degree, dimension, level, nb_distinct)


def run(hp, which: str = "oracle", threads: int | None = None, naive: bool = False, timing: bool = False):
    """Apply ``output[k] += kron(A_k) input[k]`` to a ``HostProblem`` on the CPU.

    Returns the resulting output slab (a fresh array; ``hp`` is not modified -- the input slab is
    copied first because the algorithm clobbers it, ``kronmult_omp/kronmult.hpp:47-51``).
    ``which``: ``oracle`` (C restatement), ``ref`` (reference header, -O3 + FMA contraction, the CPU
    baseline), ``ref_strict`` (reference header, -ffp-contract=off).  ``naive=True`` runs the
    explicit-Kronecker product instead (``tests/utils/kronmult_naive.h:108-121``).
    ``threads`` sets OMP_NUM_THREADS-equivalent via omp_set_num_threads (1 = deterministic sums).
    With ``timing=True`` returns ``(out, seconds)`` where seconds covers the library call only.
    """
    lib = _lib(which)
    sfx = _sfx(hp.dtype)
    ct = ctypes.c_double if sfx == "f64" else ctypes.c_float
    N = hp.N
    mats = np.ascontiguousarray(hp.mat_slab)
    inp = np.array(hp.in_slab, copy=True)
    out = np.array(hp.out_slab, copy=True)
    mo = np.ascontiguousarray(hp.mat_off, dtype=np.int64)
    io = np.ascontiguousarray(hp.in_off, dtype=np.int64)
    oo = np.ascontiguousarray(hp.out_off, dtype=np.int64)
    if threads is not None:
        try:
            gomp = ctypes.CDLL("libgomp.so.1")
            gomp.omp_set_num_threads(int(threads))
        except OSError:
            pass
    pre = _PREFIX[which]
    if naive:
        f = getattr(lib, f"{pre}_kronmult_batched_naive_slab_{sfx}")
        f.restype = None
        t0 = time.perf_counter()
        f(hp.d, hp.n, _p(mats, ct), _p(mo, ctypes.c_longlong), hp.lda, _p(inp, ct), _p(io, ctypes.c_longlong),
          _p(out, ct), _p(oo, ctypes.c_longlong), hp.nb)
        dt = time.perf_counter() - t0
    else:
        ws = np.zeros(hp.nb * N, dtype=hp.dtype)
        wo = np.arange(hp.nb, dtype=np.int64) * N
        f = getattr(lib, f"{pre}_kronmult_batched_slab_{sfx}")
        f.restype = None
        t0 = time.perf_counter()
        f(hp.d, hp.n, _p(mats, ct), _p(mo, ctypes.c_longlong), hp.lda, _p(inp, ct), _p(io, ctypes.c_longlong),
          _p(out, ct), _p(oo, ctypes.c_longlong), _p(ws, ct), _p(wo, ctypes.c_longlong), hp.nb)
        dt = time.perf_counter() - t0
    return (out, dt) if timing else out


def time_batched(hp, which: str = "ref", threads: int | None = None, reps: int = 3):
    """CPU-baseline timing helper: pointer arrays and workspaces are built and pages touched
    *outside* the timed region; returns the list of per-call seconds (call only).  Inputs are
    restored between repetitions because the algorithm clobbers them."""
    lib = _lib(which)
    sfx = _sfx(hp.dtype)
    ct = ctypes.c_double if sfx == "f64" else ctypes.c_float
    N = hp.N
    if threads is not None:
        try:
            ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(threads))
        except OSError:
            pass
    mats = np.ascontiguousarray(hp.mat_slab)
    inp = np.array(hp.in_slab, copy=True)
    out = np.array(hp.out_slab, copy=True)
    ws = np.zeros(hp.nb * N, dtype=hp.dtype)
    s = inp.itemsize
    PT = ctypes.POINTER(ct)
    A = (mats.ctypes.data + hp.mat_off.astype(np.int64) * s).astype(np.uint64)
    I = (inp.ctypes.data + hp.in_off.astype(np.int64) * s).astype(np.uint64)
    O = (out.ctypes.data + hp.out_off.astype(np.int64) * s).astype(np.uint64)
    W = (ws.ctypes.data + np.arange(hp.nb, dtype=np.int64) * N * s).astype(np.uint64)
    f = getattr(lib, f"{_PREFIX[which]}_kronmult_batched_{sfx}")
    f.restype = None
    f.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                  ctypes.c_void_p, ctypes.c_int]
    times = []
    for _ in range(reps):
        np.copyto(inp, hp.in_slab)
        t0 = time.perf_counter()
        f(hp.d, hp.n, A.ctypes.data, hp.lda, I.ctypes.data, O.ctypes.data, W.ctypes.data, hp.nb)
        times.append(time.perf_counter() - t0)
    return times


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    """Full relative L2 over every element, ``|a-b|_2 / |b|_2`` (the parity metric of
    BASELINE.json; not the reference's defective 5-element ``distance()``,
    ``tests/utils/utils_gpu.h:153-168``)."""
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    den = float(np.linalg.norm(b))
    return float(np.linalg.norm(a - b)) / (den if den > 0 else 1.0)
