"""The reference's sweep (tests/kronmult_fullbench_gpu.cpp:70-87: degree 2..10 x dimension 1..6 x level 2..9,
batch size from tests/utils/batch_size.h:8-21, matrix_stride 67, 5 distinct outputs) against this library.

    python tools/fullbench.py [--levels 9] [--dtype f64] [--ref-gpu]

One JSON line per case: kernel family, CUDA-event time of the stream-ordered C-ABI call (min of --reps), GFLOP/s,
algorithmic GB/s and the fraction of the applicable roofline; with --ref-gpu also the reference CUDA kernel
(oracle/_ref/libkronmult_refgpu.so, built for sm_100a) on the same problem.  Development / reporting tool:
bench.py is the graded benchmark.
"""
import argparse
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kronmult993_b200 import api, batch  # noqa: E402


def time_call(fn, reps, stream):
    best = float("inf")
    with torch.cuda.stream(stream):
        for r in range(reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            e1.synchronize()
            if r > 0:
                best = min(best, e0.elapsed_time(e1))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--levels", default="9", help="comma list of grid levels (reference sweeps 2..9)")
    ap.add_argument("--degrees", default="2,3,4,5,6,7,8,9,10")
    ap.add_argument("--dims", default="1,2,3,4,5,6")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--ref-gpu", action="store_true")
    ap.add_argument("--path", default="auto", help="force a kernel family (api.PATHS)")
    ap.add_argument("--tune", default="", help="knob=value[,knob=value] for kronmult_b200_set_tuning")
    ap.add_argument("--target-mb", type=float, default=0.0,
                    help="instead of the reference's batch size use a batch whose inputs total this many MB "
                         "(throughput rather than launch latency); aliasing then is runs of 32 items per output")
    ap.add_argument("--max-bytes-item", type=float, default=0.0, help="skip shapes whose vector exceeds this many bytes")
    ap.add_argument("--fp64-tflops", type=float, default=34.1)
    ap.add_argument("--fp32-tflops", type=float, default=70.8)
    args = ap.parse_args()
    dt = torch.float64 if args.dtype == "f64" else torch.float32
    hbm = 6552.0
    mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(mp):
        hbm = json.load(open(mp)).get("hbm_gbs", hbm)
    peak = (args.fp64_tflops if dt == torch.float64 else args.fp32_tflops) * 1e12
    torch.cuda.set_device(0)
    for kv in [x for x in args.tune.split(",") if x]:
        k, v = kv.split("=")
        assert api.load_library().kronmult_b200_set_tuning(int(k), int(v)) == 0
    stream = torch.cuda.Stream()
    ref = None
    if args.ref_gpu:
        ref = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libkronmult_refgpu.so"), mode=ctypes.RTLD_LOCAL)
    for level in [int(x) for x in args.levels.split(",")]:
        for d in [int(x) for x in args.dims.split(",")]:
            for n in [int(x) for x in args.degrees.split(",")]:
                nb = batch.compute_batch_size(n, d, level)
                esz = 8 if dt == torch.float64 else 4
                if args.max_bytes_item and n ** d * esz > args.max_bytes_item:
                    continue
                if args.target_mb > 0:
                    nb = max(64, int(args.target_mb * 1e6 / (n ** d * esz)))
                    p = batch.make_problem(d, n, nb, dt, "cuda", seed=993, alias="runs", items_per_output=32)
                else:
                    p = batch.make_problem(d, n, nb, dt, "cuda", seed=993, alias="ref", nb_distinct=5,
                                           matrices="reftest")
                A, i, o, w = p.pointer_arrays()
                torch.cuda.synchronize()
                api.force_path(args.path)
                try:
                    ms = time_call(lambda: api.kronmult_batched(d, n, A, p.lda, i, o, w, nb, dtype=dt, stream=stream),
                                   args.reps, stream)
                except api.KronmultError:
                    api.force_path("auto")
                    continue
                api.force_path("auto")
                fl, by = p.flops(), p.algorithmic_bytes()
                roof = max(by / (hbm * 1e9), fl / peak)
                line = {"n": n, "d": d, "level": level, "nb": nb, "N": n ** d, "path": api.last_path(),
                        "ms": round(ms, 4), "gflops": round(fl / ms * 1e-6, 1), "alg_gbs": round(by / ms * 1e-6, 1),
                        "roofline_frac": round(roof * 1e3 / ms, 4)}
                if ref is not None:
                    p.alloc_workspaces()
                    A, i, o, w = p.pointer_arrays()
                    fn = getattr(ref, f"refgpu_kronmult_batched_{args.dtype}")
                    fn.restype = ctypes.c_int
                    fn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
                    backup = p.in_slab.clone()
                    best = float("inf")
                    for r in range(2):
                        p.in_slab.copy_(backup)  # the reference clobbers its input
                        torch.cuda.synchronize()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        rc = fn(d, n, A.data_ptr(), p.lda, i.data_ptr(), o.data_ptr(), w.data_ptr(), nb)
                        e1.record(); e1.synchronize()
                        assert rc == 0
                        best = min(best, e0.elapsed_time(e1))
                    line["ref_cuda_ms"] = round(best, 4)
                    line["speedup_vs_ref_cuda"] = round(best / ms, 2)
                print(json.dumps(line), flush=True)
                del p, A, i, o, w
                torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
