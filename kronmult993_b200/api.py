"""Python mirror of the reference's ``kronmult_batched`` operator, bound to the C ABI.

The reference (project-asgard/kronmult993) is a C++ library; its operator is
``kronmult_batched<T>(matrix_count, matrix_size, matrix_list_batched, matrix_stride, input_batched,
output_batched, workspace_batched, nb_batch)`` (``kronmult_gpu/kronmult.cuh:28-32``).  This module
exposes the same operator with the same argument names, order and meaning on top of
``include/kronmult_b200.h`` via ctypes so that the parity tests read like the reference's own
(``tests/kronmult_test_gpu.cpp:43-46``).  Pointer arrays are ``torch.int64`` tensors (or raw integer
addresses) that live on the device, exactly as the reference requires (``kronmult.cuh:22``).

There is no fallback of any kind: if ``libkronmult_b200.so`` cannot be loaded the import of the
operator raises, and a CUDA error code from the library is raised as ``KronmultError`` (the reference's
harness does the same with ``checkCudaErrorCode``, ``tests/utils/utils_gpu.h:11-17``).
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Union

import torch

from . import build as _build

PATH_AUTO, PATH_GENERIC, PATH_TINY, PATH_REGTILE, PATH_DMMA, PATH_WSPEC, PATH_WSPEC5, PATH_PAIRTILE = 0, 1, 2, 3, 4, 5, 6, 7
PATHS = {"auto": PATH_AUTO, "generic": PATH_GENERIC, "tiny": PATH_TINY, "regtile": PATH_REGTILE, "dmma": PATH_DMMA,
         "wspec": PATH_WSPEC, "wspec5": PATH_WSPEC5, "pairtile": PATH_PAIRTILE}

# every symbol include/kronmult_b200.h declares (tests/test_abi.py checks the header against this)
C_SYMBOLS = (
    "kronmult_pow_int",
    "kronmult_batched_f64",
    "kronmult_batched_f32",
    "kronmult_batched_f64_async",
    "kronmult_batched_f32_async",
    "kronmult_batched_host_f64",
    "kronmult_batched_host_f32",
    "kronmult_partition_by_output",
    "kronmult_plan_create_f64",
    "kronmult_plan_create_f32",
    "kronmult_batched_const_f64",
    "kronmult_batched_const_f32",
    "kronmult_batched_const_f64_async",
    "kronmult_batched_const_f32_async",
    "kronmult_b200_needs_workspace",
    "kronmult_plan_execute",
    "kronmult_plan_stats",
    "kronmult_plan_destroy",
    "kronmult_b200_plan_cache_hits",
    "kronmult_b200_plan_cache_builds",
    "kronmult_b200_version",
    "kronmult_b200_launch_count",
    "kronmult_b200_last_path",
    "kronmult_b200_force_path",
    "kronmult_b200_set_tuning",
)
# the C++ drop-in symbols of include/kronmult.cuh (same mangling as the reference library)
CXX_SYMBOLS = (
    "_Z7pow_intii",
    "_Z16kronmult_batchedIdE9cudaErroriiPKPKT_iPPS1_S7_S7_i",
    "_Z16kronmult_batchedIfE9cudaErroriiPKPKT_iPPS1_S7_S7_i",
)


class KronmultError(RuntimeError):
    def __init__(self, code: int, where: str):
        self.code = int(code)
        super().__init__(f"{where}: CUDA error {self.code}")


_lib: Optional[ctypes.CDLL] = None


def library_path() -> str:
    return _build.LIB


def load_library() -> ctypes.CDLL:
    """Load (building first if the sources are newer) the CUDA library.  Raises if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if os.environ.get("KRONMULT_B200_LIB"):  # A/B experiments: an alternative build of the same library
        path = os.environ["KRONMULT_B200_LIB"]
    elif os.environ.get("KRONMULT_B200_NO_BUILD") != "1":
        try:
            path = _build.build_library()
        except Exception:
            if not os.path.exists(path):
                raise
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing and could not be built: kronmult993_b200 has no non-CUDA path")
    lib = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
    c_int, c_vp = ctypes.c_int, ctypes.c_void_p
    for sfx in ("f64", "f32"):
        f = getattr(lib, f"kronmult_batched_{sfx}")
        f.restype, f.argtypes = c_int, [c_int, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_int]
        f = getattr(lib, f"kronmult_batched_{sfx}_async")
        f.restype, f.argtypes = c_int, [c_int, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_vp]
        f = getattr(lib, f"kronmult_batched_host_{sfx}")
        f.restype, f.argtypes = c_int, [c_int, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_int]
        f = getattr(lib, f"kronmult_plan_create_{sfx}")
        f.restype, f.argtypes = c_int, [c_int, c_int, c_vp, c_int, c_vp, c_vp, c_int, c_vp, ctypes.POINTER(c_vp)]
    for sfx in ("f64", "f32"):
        f = getattr(lib, f"kronmult_batched_const_{sfx}")
        f.restype, f.argtypes = c_int, [c_int, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_int]
        f = getattr(lib, f"kronmult_batched_const_{sfx}_async")
        f.restype, f.argtypes = c_int, [c_int, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_vp]
    lib.kronmult_b200_needs_workspace.restype = c_int
    lib.kronmult_b200_needs_workspace.argtypes = [c_int, c_int, c_int]
    lib.kronmult_plan_execute.restype, lib.kronmult_plan_execute.argtypes = c_int, [c_vp, c_vp]
    lib.kronmult_plan_stats.restype = c_int
    lib.kronmult_plan_stats.argtypes = [c_vp, ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_longlong),
                                        ctypes.POINTER(c_int)]
    lib.kronmult_plan_destroy.restype, lib.kronmult_plan_destroy.argtypes = c_int, [c_vp]
    lib.kronmult_b200_plan_cache_hits.restype = ctypes.c_longlong
    lib.kronmult_b200_plan_cache_builds.restype = ctypes.c_longlong
    lib.kronmult_pow_int.restype, lib.kronmult_pow_int.argtypes = c_int, [c_int, c_int]
    lib.kronmult_partition_by_output.restype = c_int
    lib.kronmult_partition_by_output.argtypes = [c_vp, c_int, c_int, ctypes.c_longlong, c_vp, c_vp]
    lib.kronmult_b200_version.restype = ctypes.c_char_p
    lib.kronmult_b200_launch_count.restype = ctypes.c_longlong
    lib.kronmult_b200_last_path.restype = ctypes.c_char_p
    lib.kronmult_b200_force_path.restype, lib.kronmult_b200_force_path.argtypes = c_int, [c_int]
    _lib = lib
    return lib


def _addr(x) -> int:
    if x is None:
        return 0
    if isinstance(x, torch.Tensor):
        if x.dtype != torch.int64:
            raise TypeError("pointer arrays must be int64 tensors of device addresses")
        if not x.is_contiguous():
            raise ValueError("pointer arrays must be contiguous")
        return x.data_ptr()
    return int(x)


def _suffix(dtype) -> str:
    if dtype in (torch.float64, "f64", "double"):
        return "f64"
    if dtype in (torch.float32, "f32", "float"):
        return "f32"
    # the reference only instantiates float and double (kronmult.cu:202-224): other T = link error
    raise TypeError(f"kronmult_batched is defined for float32 and float64 only, got {dtype}")


def pow_int(number: int, power: int) -> int:
    """``pow_int`` of the reference (``kronmult_gpu/kronmult.cuh:10``)."""
    return load_library().kronmult_pow_int(int(number), int(power))


def kronmult_batched(matrix_count: int, matrix_size: int, matrix_list_batched, matrix_stride: int, input_batched,
                     output_batched, workspace_batched, nb_batch: int, *, dtype=torch.float64,
                     stream: Union[None, int, "torch.cuda.Stream"] = None) -> None:
    """``output[k] += kron(matrix_list[k]) @ input[k]`` for ``k < nb_batch`` on the current device.

    Same argument list as the reference operator.  With ``stream=None`` the call is blocking and
    runs on the legacy default stream, like ``kronmult.cu:191-196``; passing a ``torch.cuda.Stream``
    (or a raw ``cudaStream_t``) uses the stream-ordered entry point and does not synchronise.
    """
    lib = load_library()
    sfx = _suffix(dtype)
    args = [int(matrix_count), int(matrix_size), _addr(matrix_list_batched), int(matrix_stride),
            _addr(input_batched), _addr(output_batched), _addr(workspace_batched), int(nb_batch)]
    if stream is None:
        code = getattr(lib, f"kronmult_batched_{sfx}")(*args)
    else:
        handle = stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream)
        code = getattr(lib, f"kronmult_batched_{sfx}_async")(*args, handle)
    if code != 0:
        raise KronmultError(code, "kronmult_batched")


def kronmult_batched_const(matrix_count: int, matrix_size: int, matrix_list_batched, matrix_stride: int, input_batched,
                           output_batched, workspace_batched, nb_batch: int, *, dtype=torch.float64,
                           stream: Union[None, int, "torch.cuda.Stream"] = None) -> None:
    """Read-only-input variant (``kronmult_batched_const_*`` of ``include/kronmult_b200.h``): ``input[k]`` is never
    written, so entries of ``input_batched`` may repeat; ``workspace_batched`` may be ``None`` unless
    ``needs_workspace(matrix_count, matrix_size, dtype)``."""
    lib = load_library()
    sfx = _suffix(dtype)
    args = [int(matrix_count), int(matrix_size), _addr(matrix_list_batched), int(matrix_stride),
            _addr(input_batched), _addr(output_batched), _addr(workspace_batched), int(nb_batch)]
    if stream is None:
        code = getattr(lib, f"kronmult_batched_const_{sfx}")(*args)
    else:
        handle = stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream)
        code = getattr(lib, f"kronmult_batched_const_{sfx}_async")(*args, handle)
    if code != 0:
        raise KronmultError(code, "kronmult_batched_const")


def needs_workspace(matrix_count: int, matrix_size: int, dtype=torch.float64) -> bool:
    r = load_library().kronmult_b200_needs_workspace(int(matrix_count), int(matrix_size),
                                                     8 if _suffix(dtype) == "f64" else 4)
    if r < 0:
        raise ValueError("invalid shape")
    return bool(r)


def kronmult_batched_host(matrix_count: int, matrix_size: int, matrix_list_batched, matrix_stride: int,
                          input_batched, output_batched, workspace_batched, nb_batch: int, *,
                          dtype=torch.float64, device: int = -1) -> None:
    """Host-memory flavour (signature of ``kronmult_omp/kronmult.hpp:77-80``): pointer arrays are raw
    addresses of HOST arrays of HOST pointers; the library stages the data through the GPU."""
    lib = load_library()
    sfx = _suffix(dtype)
    code = getattr(lib, f"kronmult_batched_host_{sfx}")(
        int(matrix_count), int(matrix_size), _addr(matrix_list_batched), int(matrix_stride), _addr(input_batched),
        _addr(output_batched), _addr(workspace_batched), int(nb_batch), int(device))
    if code != 0:
        raise KronmultError(code, "kronmult_batched_host")


class Plan:
    """An aliasing plan (``kronmult_plan_*`` of ``include/kronmult_b200.h``): the batch sorted by output
    pointer on the device so that items sharing an output are consecutive.  The pointer-array tensors are
    kept alive by the plan; the data they point to may change between executions."""

    def __init__(self, matrix_count, matrix_size, matrix_list_batched, matrix_stride, input_batched, output_batched,
                 nb_batch, *, dtype=torch.float64, stream=None):
        lib = load_library()
        self._keep = (matrix_list_batched, input_batched, output_batched)
        self._h = ctypes.c_void_p()
        handle = 0 if stream is None else (stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream))
        code = getattr(lib, f"kronmult_plan_create_{_suffix(dtype)}")(
            int(matrix_count), int(matrix_size), _addr(matrix_list_batched), int(matrix_stride), _addr(input_batched),
            _addr(output_batched), int(nb_batch), handle, ctypes.byref(self._h))
        if code != 0:
            raise KronmultError(code, "kronmult_plan_create")

    def execute(self, stream=None) -> None:
        handle = 0 if stream is None else (stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream))
        code = load_library().kronmult_plan_execute(self._h, handle)
        if code != 0:
            raise KronmultError(code, "kronmult_plan_execute")

    def stats(self) -> dict:
        rb, ra, pm = ctypes.c_longlong(), ctypes.c_longlong(), ctypes.c_int()
        load_library().kronmult_plan_stats(self._h, ctypes.byref(rb), ctypes.byref(ra), ctypes.byref(pm))
        return {"runs_before": rb.value, "runs_after": ra.value, "permuted": bool(pm.value)}

    def destroy(self) -> None:
        if self._h:
            load_library().kronmult_plan_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def plan_cache_counters() -> "tuple[int, int]":
    """(hits, builds) of the implicit plans of the blocking entry points."""
    lib = load_library()
    return int(lib.kronmult_b200_plan_cache_hits()), int(lib.kronmult_b200_plan_cache_builds())


def run_problem(problem, *, stream=None, path: str = "auto") -> None:
    """Apply the operator to a ``batch.KronProblem`` in place (``problem.out_slab`` accumulates)."""
    A, i, o, w = problem.pointer_arrays()
    force_path(path)
    try:
        kronmult_batched(problem.d, problem.n, A, problem.lda, i, o, w, problem.nb, dtype=problem.dtype,
                         stream=stream)
    finally:
        force_path("auto")


def force_path(path: Union[str, int]) -> None:
    code = load_library().kronmult_b200_force_path(PATHS[path] if isinstance(path, str) else int(path))
    if code != 0:
        raise KronmultError(code, "kronmult_b200_force_path")


def set_tuning(knob: int, value: int) -> None:
    code = load_library().kronmult_b200_set_tuning(int(knob), int(value))
    if code != 0:
        raise KronmultError(code, "kronmult_b200_set_tuning")


def last_path() -> str:
    return load_library().kronmult_b200_last_path().decode()


def launch_count() -> int:
    return int(load_library().kronmult_b200_launch_count())


def version() -> str:
    return load_library().kronmult_b200_version().decode()
