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

PATH_AUTO, PATH_GENERIC, PATH_TINY, PATH_REGTILE, PATH_DMMA, PATH_WSPEC, PATH_WSPEC5, PATH_PAIRTILE, PATH_SYM5, PATH_SYM4 = 0, 1, 2, 3, 4, 5, 6, 7, 8, 9
PATHS = {"auto": PATH_AUTO, "generic": PATH_GENERIC, "tiny": PATH_TINY, "regtile": PATH_REGTILE, "dmma": PATH_DMMA,
         "wspec": PATH_WSPEC, "wspec5": PATH_WSPEC5, "pairtile": PATH_PAIRTILE, "sym5": PATH_SYM5, "sym4": PATH_SYM4}

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
    "kronmult_comm_unique_id",
    "kronmult_comm_create",
    "kronmult_comm_adopt",
    "kronmult_comm_destroy",
    "kronmult_comm_last_collective_ms",
    "kronmult_batched_sharded_f64",
    "kronmult_batched_sharded_f32",
    "kronmult_build_batch_f64",
    "kronmult_build_batch_f32",
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
    for sfx in ("f64", "f32"):
        f = getattr(lib, f"kronmult_batched_sharded_{sfx}")
        f.restype = c_int
        f.argtypes = [c_int, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_vp]
    for sfx in ("f64", "f32"):
        f = getattr(lib, f"kronmult_build_batch_{sfx}")
        f.restype = c_int
        f.argtypes = [c_int, c_int, c_int, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp,
                      ctypes.POINTER(ctypes.c_longlong), c_vp]
    lib.kronmult_comm_unique_id.restype, lib.kronmult_comm_unique_id.argtypes = c_int, [c_vp]
    lib.kronmult_comm_create.restype = c_int
    lib.kronmult_comm_create.argtypes = [c_vp, c_int, c_int, ctypes.POINTER(c_vp)]
    lib.kronmult_comm_adopt.restype = c_int
    lib.kronmult_comm_adopt.argtypes = [c_vp, c_int, c_int, ctypes.POINTER(c_vp)]
    lib.kronmult_comm_destroy.restype, lib.kronmult_comm_destroy.argtypes = c_int, [c_vp]
    lib.kronmult_comm_last_collective_ms.restype = c_int
    lib.kronmult_comm_last_collective_ms.argtypes = [c_vp, ctypes.POINTER(ctypes.c_float),
                                                     ctypes.POINTER(ctypes.c_longlong)]
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


def build_asgard_batch(d: int, n: int, lda: int, cells: torch.Tensor, coeff: torch.Tensor, nterms: int, rows, cols,
                       x: torch.Tensor, y: torch.Tensor, *, stream=None):
    """Device-side pointer arrays of the ASGarD batch {(i, j, t)} (``kronmult_build_batch_*``).

    ``cells`` int32 [num_elements, d] 1-D cell indices, ``coeff`` int64 [nterms*d] device addresses of the 1-D
    coefficient matrices, ``rows`` / ``cols`` = (first, last+1) element ranges, ``x`` / ``y`` the stacked element
    vectors.  Returns ``(A, in, out, nb)`` as int64 device tensors for ``kronmult_batched_const``."""
    lib = load_library()
    assert cells.dtype == torch.int32 and cells.is_contiguous() and coeff.dtype == torch.int64
    nb = (rows[1] - rows[0]) * (cols[1] - cols[0]) * nterms
    dev = x.device
    A = torch.empty(max(1, nb * d), dtype=torch.int64, device=dev)
    i_ = torch.empty(max(1, nb), dtype=torch.int64, device=dev)
    o_ = torch.empty(max(1, nb), dtype=torch.int64, device=dev)
    got = ctypes.c_longlong()
    handle = 0 if stream is None else (stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream))
    code = getattr(lib, f"kronmult_build_batch_{_suffix(x.dtype)}")(
        int(d), int(n), int(lda), cells.data_ptr(), coeff.data_ptr(), int(nterms), int(rows[0]), int(rows[1]),
        int(cols[0]), int(cols[1]), x.data_ptr(), y.data_ptr(), A.data_ptr(), i_.data_ptr(), o_.data_ptr(),
        ctypes.byref(got), handle)
    if code != 0:
        raise KronmultError(code, "kronmult_build_batch")
    assert got.value == nb
    return A[: nb * d], i_[:nb], o_[:nb], nb


class Comm:
    """A ``kronmult_comm`` over the ranks of a ``torch.distributed`` job (one process per GPU): rank 0's
    ncclUniqueId is broadcast with ``torch.distributed`` (plumbing), the communicator itself is the library's own."""

    def __init__(self, dist=None, device=None):
        import numpy as np
        lib = load_library()
        self.world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self._h = ctypes.c_void_p()
        ident = np.zeros(128, dtype=np.uint8)
        if self.world > 1:
            if self.rank == 0:
                code = lib.kronmult_comm_unique_id(ident.ctypes.data)
                if code != 0:
                    raise KronmultError(code, "kronmult_comm_unique_id")
            t = torch.from_numpy(ident)
            if device is not None:
                t = t.to(device)
            dist.broadcast(t, 0)
            ident = t.cpu().numpy().copy()
        code = lib.kronmult_comm_create(ident.ctypes.data, self.world, self.rank, ctypes.byref(self._h))
        if code != 0:
            raise KronmultError(code, "kronmult_comm_create")

    def last_collective(self) -> "tuple[float, int]":
        ms, cnt = ctypes.c_float(), ctypes.c_longlong()
        code = load_library().kronmult_comm_last_collective_ms(self._h, ctypes.byref(ms), ctypes.byref(cnt))
        if code != 0:
            raise KronmultError(code, "kronmult_comm_last_collective_ms")
        return float(ms.value), int(cnt.value)

    def destroy(self) -> None:
        if self._h:
            load_library().kronmult_comm_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def kronmult_batched_sharded(matrix_count: int, matrix_size: int, matrix_list_batched, matrix_stride: int,
                             input_batched, output_batched, workspace_batched, nb_batch: int, shared_outputs,
                             comm: "Comm", *, owner=None, dtype=torch.float64, stream=None) -> None:
    """This rank's shard of a batch (``kronmult_batched_sharded_*``): ``shared_outputs`` lists the device
    addresses of this rank's copies of the output vectors that were split across ranks (same order on every
    rank); their partial sums are combined with one NCCL all-reduce inside the call."""
    import numpy as np
    lib = load_library()
    sh = np.ascontiguousarray(np.asarray(shared_outputs, dtype=np.uint64))
    ow = None if owner is None else np.ascontiguousarray(np.asarray(owner, dtype=np.int32))
    handle = 0 if stream is None else (stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream))
    code = getattr(lib, f"kronmult_batched_sharded_{_suffix(dtype)}")(
        int(matrix_count), int(matrix_size), _addr(matrix_list_batched), int(matrix_stride), _addr(input_batched),
        _addr(output_batched), _addr(workspace_batched), int(nb_batch), sh.ctypes.data if sh.size else 0, int(sh.size),
        ow.ctypes.data if ow is not None else 0, comm._h, handle)
    if code != 0:
        raise KronmultError(code, "kronmult_batched_sharded")


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
