#!/usr/bin/env python
"""bench.py -- the graded benchmark of the kronmult_batched hot path.

    python bench.py --gpus N --steps K --warmup W          (N>1: launched by torch.distributed.run)
    python bench.py --impl reference ...                   (the reference's own CPU path, same metric)

A "step" is ONE call of kronmult_batched over the whole batch of the workload.  Default workload is
BASELINE.json's multi-GPU configuration, C5-f64: d=5 factors of n=4 (N=1024-element vectors),
8 Mi batch items, ASGarD-style aliasing (32 consecutive items per output, 262144 outputs), sharded
over the GPUs by output-pointer ownership (no data-path collective; total work is fixed, so
"scaling": "strong").  Other BASELINE configurations: --config c2|c3|c4a|c4b|c5_f32|c1.

value    whole-job GFLOP/s (FLOPs = nb*2*d*n^(d+1)) with all operands resident in HBM, timed with CUDA
         events on the launching stream, barrier + synchronize on both sides, MAX over ranks.
roofline dominant (only) kernel: algorithmic bytes per launch / measured launch time vs the MEASURED
         HBM copy bandwidth (MEASURED_PEAKS.json); the FP-pipe bound is reported next to it.
e2e      same metric through the host-buffer C ABI (kronmult_batched_host_*): pinned host memory,
         H2D of inputs/factors/outputs and D2H of the outputs inside the timed region.
cpu_baseline  the reference's kronmult_omp (oracle/_ref, compiled from /root/reference where it lies;
         else the C port in oracle/) on the host cores, on a bounded sample of the same workload.
configs  (N=1, or --all-configs) every other BASELINE.json configuration at full size -- C2, C3, C4a, C4b, C5-f32,
         C1: sustained and best ms, GFLOP/s, algorithmic GB/s, fraction of the applicable roofline, kernel family,
         and the relative L2 error of ONE application against the CPU oracle on a strided subset of output groups.
blocking the real drop-in call kronmult_batched<double> (legacy stream, implicit plan scan, device sync) on the
         headline workload: first (cold) call and plan-cached calls.
asgard_layout  the headline shape with ASGarD-style factors: 4x4 windows into 5 shared 256x256 coefficient
         matrices, lda = 256 (20 strided 32-byte column copies per item instead of one 640-byte copy).
reference_gpu  the reference's kronmult_gpu/kronmult.cu rebuilt for sm_100a (oracle/_ref), same box, same run,
         on 1/8 of the headline batch (it needs a real workspace per item and restores its clobbered inputs).
ref5     the reference harness' own aliasing pattern -- 5 distinct outputs (tests/kronmult_bench_gpu.cpp:15) on
         its `realistic` case (n=8, d=6, 3903 items of 262144 elements, stride 67) -- sharded over the N ranks
         with every output group split: kronmult_batched_sharded_* with ONE ncclAllReduce (5 x 2 MiB) on the
         timed path; reports the collective's device time and the parity of a small instance vs the oracle.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fp64 batched kronmult GFLOP/s + effective HBM GB/s vs B200 roofline, 1/2/4/8 GPUs"

# name: (d, n, nb, dtype, items_per_output)
CONFIGS = {
    "c1": (3, 4, 65536, "f64", 1),
    "c2": (2, 2, 1 << 24, "f64", 1),
    "c3": (6, 4, 1 << 20, "f64", 32),
    "c4a": (4, 8, 1 << 19, "f64", 1),
    "c4b": (4, 8, 1 << 19, "f64", 32),
    "c5_f64": (5, 4, 1 << 23, "f64", 32),
    "c5_f32": (5, 4, 1 << 23, "f32", 32),
}
# measured on this pool's B200 by kron_microbench (profiles/microbench_r01.jsonl)
FP_PEAK_TFLOPS = {"f64": 34.1, "f32": 70.8}


def workload_name(cfg, d, n, nb, dt, r):
    return (f"{cfg}: {dt} d={d} n={n} (N={n ** d}) nb={nb} items, "
            + (f"{r} items/output ({nb // r} outputs, ASGarD-style runs)" if r > 1 else "distinct outputs"))


def measured_hbm_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def host_sample_problem(d, n, items, dt, r, seed=993):
    import torch
    from kronmult993_b200 import batch
    tdt = torch.float64 if dt == "f64" else torch.float32
    return batch.make_problem(d, n, items, tdt, "cpu", seed=seed, alias="runs" if r > 1 else "distinct",
                              items_per_output=r).to_host()


def cpu_reference_gflops(d, n, dt, r, sample_items, reps):
    """The reference's kronmult_omp on the host cores, all threads, call-only timing."""
    from oracle import oracle
    which = "ref" if oracle.available("ref") else "oracle"
    hp = host_sample_problem(d, n, sample_items, dt, r)
    cores = os.cpu_count() or 1
    times = oracle.time_batched(hp, which, threads=cores, reps=reps)
    flops = sample_items * 2 * d * n ** (d + 1)
    return flops / min(times) * 1e-9, flops, times, cores, ("reference" if which == "ref" else "port")


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path, same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    d, n, nb, dt, r = CONFIGS[args.config]
    sample = min(nb, args.cpu_sample)
    from oracle import oracle
    which = "ref" if oracle.available("ref") else "oracle"
    hp = host_sample_problem(d, n, sample, dt, r)
    cores = os.cpu_count() or 1
    oracle.time_batched(hp, which, threads=cores, reps=max(1, args.warmup))
    t0 = time.perf_counter()
    times = oracle.time_batched(hp, which, threads=cores, reps=args.steps)
    wall = time.perf_counter() - t0
    flops = sample * 2 * d * n ** (d + 1)
    per_step = sum(times) / len(times)
    value = flops / per_step * 1e-9
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(per_step * 1e3, 3),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": dt, "data": "synthetic",
        "config": {"workload": workload_name(args.config, d, n, nb, dt, r),
                   "sample": f"each step = kronmult_omp over {sample} of the {nb} items (rate metric)"},
        "cpu_baseline": {"value": round(value, 3), "unit": "GFLOP/s", "cores": cores,
                         "kind": "reference" if which == "ref" else "port",
                         "sample": f"{sample} items/step, {args.steps} steps, OpenMP {cores} threads, "
                                   f"kronmult_omp no-BLAS -O3 -march=x86-64-v3; wall {wall:.1f}s"},
        "e2e": {"value": round(value, 3), "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_reference_gpu_arm(args):
    """--impl reference_gpu: the reference CUDA kernel recompiled for sm_100a (timing baseline #2)."""
    import torch
    from kronmult993_b200 import batch
    d, n, nb, dt, r = CONFIGS[args.config]
    nb = max(1, int(nb * min(args.scale, 0.125)))
    lib_path = os.path.join(ROOT, "oracle", "_ref", "libkronmult_refgpu.so")
    if not os.path.exists(lib_path):
        print(json.dumps({"impl": "reference_gpu", "unavailable": "oracle/_ref/libkronmult_refgpu.so not built"}))
        return
    lib = ctypes.CDLL(lib_path, mode=ctypes.RTLD_LOCAL)
    tdt = torch.float64 if dt == "f64" else torch.float32
    torch.cuda.set_device(0)
    p = batch.make_problem(d, n, nb, tdt, "cuda", seed=993, alias="runs" if r > 1 else "distinct", items_per_output=r)
    p.alloc_workspaces()
    A, i, o, w = p.pointer_arrays()
    fn = getattr(lib, f"refgpu_kronmult_batched_{dt}")
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                   ctypes.c_void_p, ctypes.c_int]
    backup = p.in_slab.clone()
    torch.cuda.synchronize()
    times = []
    for s in range(args.warmup + args.steps):
        p.in_slab.copy_(backup)  # the reference clobbers its input (kronmult.cu:115-121)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(d, n, A.data_ptr(), p.lda, i.data_ptr(), o.data_ptr(), w.data_ptr(), nb)
        e1.record(); e1.synchronize()
        assert rc == 0, rc
        if s >= args.warmup:
            times.append(e0.elapsed_time(e1) * 1e-3)
    per = sum(times) / len(times)
    print(json.dumps({"impl": "reference_gpu", "metric": METRIC, "value": round(p.flops() / per * 1e-9, 2),
                      "unit": "GFLOP/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": round(per * 1e3, 4), "higher_is_better": True, "dtype": dt, "data": "synthetic",
                      "config": {"workload": workload_name(args.config, d, n, nb, dt, r),
                                 "note": "reference kronmult_gpu/kronmult.cu built -arch=sm_100a, real workspaces, "
                                         "input restored between steps, blocking call incl. cudaDeviceSynchronize"},
                      "alg_gbs": round(p.algorithmic_bytes() / per * 1e-9, 1)}), flush=True)


def _tdt(dt):
    import torch
    return torch.float64 if dt == "f64" else torch.float32


def _time_calls(fn, stream, warmup, steps):
    """CUDA-event time of `steps` back-to-back calls (sustained) and of the best single call among them."""
    import torch
    with torch.cuda.stream(stream):
        for _ in range(warmup):
            fn()
        stream.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record(stream)
        for i in range(steps):
            fn()
            ev[i + 1].record(stream)
        stream.synchronize()
    per = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    return ev[0].elapsed_time(ev[steps]) / steps, min(per)


def _roof_ms(p, dt, hbm):
    return max(p.algorithmic_bytes() / (hbm * 1e9), p.flops() / (FP_PEAK_TFLOPS[dt] * 1e12)) * 1e3


def _subset_parity(p, tdt, n_groups=48):
    """relative L2 of ONE application vs the CPU oracle on a strided subset of output groups.
    Must run before anything else touches p.out_slab; leaves the outputs one application further."""
    import numpy as np
    import torch
    from kronmult993_b200 import api
    from oracle import oracle
    G = int(p.out_slab.numel() // p.N)
    step = max(1, G // n_groups)
    groups = torch.arange(0, G, step, dtype=torch.int64)[:n_groups]
    hp, groups = p.select_outputs_to_host(groups)
    A, i_, o_, w_ = p.pointer_arrays()
    api.kronmult_batched(p.d, p.n, A, p.lda, i_, o_, w_, p.nb, dtype=tdt)
    idx = (groups.to(p.device)[:, None] * p.N + torch.arange(p.N, device=p.device)[None, :]).flatten()
    got = p.out_slab[idx].cpu().numpy()
    exp = oracle.run(hp, "oracle", threads=os.cpu_count() or 1)
    return float(oracle.rel_l2(got, exp)), int(hp.nb)


def measure_config(name, hbm, steps, warmup, device, matrices="dense"):
    import torch
    from kronmult993_b200 import api, batch
    d, n, nb, dt, r = CONFIGS[name]
    tdt = _tdt(dt)
    p = batch.make_problem(d, n, nb, tdt, device, seed=993, alias="runs" if r > 1 else "distinct", items_per_output=r,
                           matrices=matrices)
    torch.cuda.synchronize()
    try:
        err, n_checked = _subset_parity(p, tdt)
    except Exception as ex:
        err, n_checked = None, f"{type(ex).__name__}: {ex}"
    A, i_, o_, w_ = p.pointer_arrays()
    stream = torch.cuda.Stream()
    torch.cuda.synchronize()
    time.sleep(1.0)  # every configuration starts from an idle GPU (the previous one leaves it power-capped)
    ms, ms_min = _time_calls(lambda: api.kronmult_batched(p.d, p.n, A, p.lda, i_, o_, w_, p.nb, dtype=tdt, stream=stream),
                             stream, warmup, steps)
    roof = _roof_ms(p, dt, hbm)
    bound = "hbm" if p.algorithmic_bytes() / (hbm * 1e9) >= p.flops() / (FP_PEAK_TFLOPS[dt] * 1e12) else "fp"
    res = {"config": name, "workload": workload_name(name, d, n, nb, dt, r) + (f", {matrices} factors lda={p.lda}" if matrices != "dense" else ""),
           "path": api.last_path(), "ms": round(ms, 4), "ms_min": round(ms_min, 4),
           "gflops": round(p.flops() / ms * 1e-6, 1), "alg_gbs": round(p.algorithmic_bytes() / ms * 1e-6, 1),
           "bound": bound, "roofline_ms": round(roof, 4), "frac": round(roof / ms, 4), "frac_best": round(roof / ms_min, 4),
           "rel_l2_vs_oracle": err, "parity_items": n_checked, "tol": 1e-12 if dt == "f64" else 1e-5}
    del p, A, i_, o_, w_
    torch.cuda.empty_cache()
    return res


def measure_blocking(p, tdt, A, i_, o_, w_, reps=5):
    """The reference-facing drop-in call itself: kronmult_batched<T> = legacy default stream + the implicit plan
    scan (1 kernel + a 32-byte D2H) + cudaDeviceSynchronize (kronmult.cu:191-196).  Wall clock around the call."""
    import torch
    from kronmult993_b200 import api
    torch.cuda.synchronize()
    h0, b0 = api.plan_cache_counters()
    ts = []
    for _ in range(reps + 1):
        t0 = time.perf_counter()
        api.kronmult_batched(p.d, p.n, A, p.lda, i_, o_, w_, p.nb, dtype=tdt)  # stream=None: the blocking entry
        ts.append((time.perf_counter() - t0) * 1e3)
    h1, b1 = api.plan_cache_counters()
    return {"cold_ms": round(ts[0], 3), "plan_cached_ms": round(sum(ts[1:]) / reps, 3), "plan_cached_ms_min": round(min(ts[1:]), 3),
            "plan_cache_hits": h1 - h0, "plans_built": b1 - b0,
            "note": "kronmult_batched_f64/_f32 (what kronmult_batched<T> of include/kronmult.cuh forwards to): "
                    "host wall clock incl. the per-call run-count scan and cudaDeviceSynchronize"}


def measure_reference_gpu(name, steps, device):
    """The reference CUDA kernel rebuilt for sm_100a on 1/8 of the batch (it needs nb real workspaces)."""
    import torch
    from kronmult993_b200 import batch
    lib_path = os.path.join(ROOT, "oracle", "_ref", "libkronmult_refgpu.so")
    if not os.path.exists(lib_path):
        return {"unavailable": "oracle/_ref/libkronmult_refgpu.so not built"}
    d, n, nb, dt, r = CONFIGS[name]
    nb //= 8
    lib = ctypes.CDLL(lib_path, mode=ctypes.RTLD_LOCAL)
    tdt = _tdt(dt)
    p = batch.make_problem(d, n, nb, tdt, device, seed=993, alias="runs" if r > 1 else "distinct", items_per_output=r)
    p.alloc_workspaces()
    A, i_, o_, w_ = p.pointer_arrays()
    fn = getattr(lib, f"refgpu_kronmult_batched_{dt}")
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                   ctypes.c_void_p, ctypes.c_int]
    backup = p.in_slab.clone()
    times = []
    for s_ in range(1 + steps):
        p.in_slab.copy_(backup)  # the reference clobbers its input (kronmult.cu:115-121)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(d, n, A.data_ptr(), p.lda, i_.data_ptr(), o_.data_ptr(), w_.data_ptr(), nb)
        e1.record(); e1.synchronize()
        if rc != 0:
            return {"unavailable": f"reference kernel returned CUDA error {rc}"}
        if s_ >= 1:
            times.append(e0.elapsed_time(e1))
    ms = sum(times) / len(times)
    res = {"impl": "kronmult_gpu/kronmult.cu of the reference, built -arch=sm_100a (oracle/_ref/libkronmult_refgpu.so)",
           "workload": workload_name(name, d, n, nb, dt, r) + " (1/8 of the batch)", "ms": round(ms, 3),
           "gflops": round(p.flops() / ms * 1e-6, 1), "alg_gbs": round(p.algorithmic_bytes() / ms * 1e-6, 1), "steps": steps}
    del p, backup, A, i_, o_, w_
    torch.cuda.empty_cache()
    return res


def measure_ref5(world, rank, local, dist, steps, warmup):
    """The reference harness' 5-distinct-outputs pattern on its `realistic` case, every output group split over the
    ranks: kronmult_batched_sharded_* = shard kernel(s) + ONE ncclAllReduce + the owners' add, all on one stream."""
    import numpy as np
    import torch
    from kronmult993_b200 import api, batch, partition
    dev = torch.device("cuda", local)
    comm = api.Comm(dist if world > 1 else None, device=dev)
    out = {}
    # ---- parity on the reference's `medium` case (n=6, d=3, 384 items, 5 outputs) against the oracle
    try:
        from oracle import oracle
        full = batch.reference_case("medium", torch.float64, "cpu", seed=77).to_host()
        shard, local_out, split_owner = partition.run_shard_on_device(full, rank, world, comm, dev, split_threshold=1)
        N = full.N
        nw = shard.whole_keys.size
        merged = torch.zeros(full.out_slab.size, dtype=torch.float64, device=dev)
        for j, key in enumerate(shard.split_keys):
            if split_owner[j] == rank:
                merged[int(key): int(key) + N] = local_out[(nw + j) * N: (nw + j + 1) * N]
        for j, key in enumerate(shard.whole_keys):
            merged[int(key): int(key) + N] = local_out[j * N: (j + 1) * N]
        if world > 1:
            dist.all_reduce(merged)
        if rank == 0:
            exp = oracle.run(full, "oracle", threads=1)
            out["rel_l2_vs_oracle"] = float(oracle.rel_l2(merged.cpu().numpy(), exp))
            out["parity_case"] = f"reference `medium` (n=6 d=3 nb={full.nb}, stride 67, 5 outputs), {int(shard.split_keys.size)} split groups"
    except Exception as ex:
        out["rel_l2_vs_oracle"] = None
        out["parity_error"] = f"{type(ex).__name__}: {ex}"
    # ---- timing on `realistic`: this rank's slice of every output group
    n, d, level = batch.REFERENCE_CASES["realistic"]
    nb_total = batch.compute_batch_size(n, d, level, 5)
    g_full, n_out = batch.output_groups(nb_total, "ref", nb_distinct=5)
    owner, red = partition.partition_by_output(g_full.numpy(), world, split_threshold=1)
    mine = np.nonzero(owner == rank)[0]
    p = batch.make_problem(d, n, int(mine.size), torch.float64, dev, seed=993 + rank, alias="ref", nb_distinct=5,
                           matrices="reftest")
    N = p.N
    p.out_off = torch.from_numpy(g_full.numpy()[mine] * N).to(dev)  # the groups of MY items
    p.out_slab = torch.randn(n_out * N, dtype=torch.float64, device=dev)
    A, i_, o_, w_ = p.pointer_arrays()
    shared = [p.out_slab.data_ptr() + j * N * 8 for j in range(n_out)] if world > 1 and int(red.sum()) > 0 else []
    stream = torch.cuda.Stream()
    torch.cuda.synchronize()

    def call():
        api.kronmult_batched_sharded(d, n, A, p.lda, i_, o_, w_, p.nb, shared, comm, owner=None, dtype=torch.float64,
                                     stream=stream)
    coll = []
    with torch.cuda.stream(stream):
        for _ in range(warmup):
            call()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            call()
        e1.record(stream)
        torch.cuda.synchronize()
        ms_c, ncoll = comm.last_collective()
    t = torch.tensor([e0.elapsed_time(e1) / steps, ms_c], dtype=torch.float64, device=dev)
    fl = torch.tensor([float(p.flops())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(fl)
    ms = float(t[0].item())
    out.update({"workload": f"reference `realistic`: f64 d={d} n={n} (N={N}) nb={nb_total} items, stride 67, 5 distinct outputs "
                            f"(tests/kronmult_bench_gpu.cpp:15,72), every group split over {world} rank(s)",
                "ms_per_step": round(ms, 4), "gflops": round(float(fl.item()) / ms * 1e-6, 1),
                "collective": (f"ncclAllReduce of {n_out} x {N * 8 >> 10} KiB partial vectors, in place, on the timed stream"
                               if shared else "none (one rank owns every output)"),
                "collective_ms_last": round(float(t[1].item()), 4), "collectives_issued": int(ncoll),
                "path": api.last_path(), "steps": steps})
    comm.destroy()
    del p, A, i_, o_, w_
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference_gpu"])
    ap.add_argument("--config", default="c5_f64", choices=list(CONFIGS))
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the batch (debugging only)")
    ap.add_argument("--cpu-sample", type=int, default=65536, help="items per CPU-baseline step")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--all-configs", action="store_true", help="measure the `configs` block at N > 1 too (rank 0 only)")
    ap.add_argument("--no-extras", action="store_true", help="headline only: skip configs / blocking / asgard / reference_gpu / ref5")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        return run_reference_arm(args)
    if args.impl == "reference_gpu":
        return run_reference_gpu_arm(args)

    # stdout carries exactly ONE line (the JSON record): NCCL prints its version banner to the C-level stdout at the
    # first communicator creation, so everything else in this process goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    import torch.distributed as dist
    from kronmult993_b200 import api, batch, partition

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    api.load_library()

    d, n, nb_total, dt, r = CONFIGS[args.config]
    nb_total = max(r, int(nb_total * args.scale) // r * r)
    tdt = torch.float64 if dt == "f64" else torch.float32
    s_el = 8 if dt == "f64" else 4
    N = n ** d

    # ---- shard by output ownership: every output vector (and all items feeding it) has one rank
    keys = np.arange(nb_total, dtype=np.int64) // r
    owner, needs_reduce = partition.partition_by_output(keys, world)
    assert int(needs_reduce.sum()) == 0, "this aliasing pattern splits cleanly: no collective on the data path"
    nb = int((owner == rank).sum())
    del keys, owner, needs_reduce
    p = batch.make_problem(d, n, nb, tdt, f"cuda:{local}", seed=993 + rank, alias="runs" if r > 1 else "distinct",
                           items_per_output=r)
    A, i_, o_, w_ = p.pointer_arrays()
    flops_rank, bytes_rank = p.flops(), p.algorithmic_bytes()
    stream = torch.cuda.Stream()
    torch.cuda.synchronize()

    def step():
        api.kronmult_batched(p.d, p.n, A, p.lda, i_, o_, w_, p.nb, dtype=tdt, stream=stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step()
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        launches0 = api.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        barrier()
        launches = api.launch_count() - launches0
        clocks = sampler.stop() if rank == 0 else None
    ms_rank = e0.elapsed_time(e1)
    path = api.last_path()
    tmax = torch.tensor([ms_rank], dtype=torch.float64, device=f"cuda:{local}")
    tot = torch.tensor([float(flops_rank), float(bytes_rank), float(launches)], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    ms_total = float(tmax.item())
    flops_total, bytes_total, launches_total = (float(x) for x in tot.tolist())
    sec_step = ms_total * 1e-3 / args.steps
    value = flops_total / sec_step * 1e-9

    traffic = ncu_traffic(args.config, p)
    # ---- end-to-end through the host-buffer C ABI (pinned host memory, copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        e2e = measure_e2e(args, p, tdt, dt, world, rank, local, dist)

    # ---- the multi-GPU fallback collective on the reference harness' own aliasing pattern (every rank takes part)
    extras = {}
    if not args.no_extras:
        try:
            extras["ref5"] = measure_ref5(world, rank, local, dist, steps=10, warmup=3)
        except Exception as ex:
            extras["ref5"] = {"error": f"{type(ex).__name__}: {ex}"}
    # ---- rank 0 only: the drop-in blocking call, the ASGarD factor layout, the other BASELINE configurations and
    # the reference CUDA kernel -- after the headline so that they cannot disturb it
    if rank == 0 and not args.no_extras:
        hbm_x, _ = measured_hbm_gbs()
        try:
            extras["blocking"] = measure_blocking(p, tdt, A, i_, o_, w_)
        except Exception as ex:
            extras["blocking"] = {"error": f"{type(ex).__name__}: {ex}"}
    if not args.no_extras:
        del p, A, i_, o_, w_
        torch.cuda.empty_cache()
    if rank == 0 and not args.no_extras and args.scale == 1.0:
        dev_s = f"cuda:{local}"
        try:
            extras["asgard_layout"] = measure_config(args.config, hbm_x, 10, 3, dev_s, matrices="asgard")
        except Exception as ex:
            extras["asgard_layout"] = {"error": f"{type(ex).__name__}: {ex}"}
        if world == 1 or args.all_configs:
            cfgs = []
            for name in ("c2", "c3", "c4a", "c4b", "c5_f32", "c5_f64", "c1"):
                if name == args.config:
                    continue
                try:
                    cfgs.append(measure_config(name, hbm_x, 10, 3, dev_s))
                except Exception as ex:
                    cfgs.append({"config": name, "error": f"{type(ex).__name__}: {ex}"})
            extras["configs"] = cfgs
            try:
                extras["reference_gpu"] = measure_reference_gpu(args.config, 2, dev_s)
            except Exception as ex:
                extras["reference_gpu"] = {"error": f"{type(ex).__name__}: {ex}"}
    if world > 1:
        dist.barrier()

    if rank == 0:
        hbm, hbm_src = measured_hbm_gbs()
        launch_s = (ms_rank * 1e-3) / args.steps
        achieved = bytes_rank / launch_s * 1e-9
        fp_peak = FP_PEAK_TFLOPS[dt]
        roofline = {
            "bound": "hbm", "achieved": round(achieved, 1), "peak": hbm, "unit": "GB/s",
            "frac": round(achieved / hbm, 4), "traffic": traffic,
            "traffic_source": "static: dram__bytes_read+write per item of the committed ncu --set full capture of this "
                              "kernel (profiles/ncu_traffic.json) x items per launch -- not measured in this run",
            "peak_source": f"{hbm_src} (MEASURED_PEAKS.json copy bandwidth)",
            "kernel": path, "launch_ms": round(launch_s * 1e3, 4),
            "alg_bytes_per_launch": bytes_rank,
            "fp_bound": {"achieved_tflops": round(flops_rank / launch_s * 1e-12, 2), "peak_tflops": fp_peak,
                         "frac": round(flops_rank / launch_s * 1e-12 / fp_peak, 4),
                         "peak_source": "measured DFMA/FFMA saturation kernel (profiles/microbench_r01.jsonl)"},
        }
        roofline["applicable_frac"] = round(max(bytes_rank / (hbm * 1e9), flops_rank / (fp_peak * 1e12)) / launch_s, 4)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                g, fl, times, cores, kind = cpu_reference_gflops(d, n, dt, r, min(nb_total, args.cpu_sample), reps=5)
                cpu = {"value": round(g, 3), "unit": "GFLOP/s", "cores": cores, "kind": kind,
                       "sample": f"{min(nb_total, args.cpu_sample)} of {nb_total} items, best of 5 calls "
                                 f"({min(times) * 1e3:.1f} ms), kronmult_omp no-BLAS, OpenMP {cores} threads"}
            except Exception as ex:  # the checker is optional for the measurement itself
                cpu = {"value": None, "unit": "GFLOP/s", "cores": os.cpu_count(), "kind": "unavailable",
                       "sample": f"{type(ex).__name__}: {ex}"}
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(sec_step * 1e3, 4), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": dt, "data": "synthetic",
            "config": {"workload": workload_name(args.config, d, n, nb_total, dt, r),
                       "sharding": f"{world} rank(s), items partitioned by output-pointer owner, no collective",
                       "l2": f"per-rank inputs {nb * N * s_el / 2**30:.1f} GiB >> 126 MB L2 (no flush needed)",
                       "kernel_path": path},
            "effective_hbm_gbs": round(bytes_total / sec_step * 1e-9, 1),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches_total),
            "clocks": clocks,
        }
        ref_gpu = extras.get("reference_gpu")
        if isinstance(ref_gpu, dict) and ref_gpu.get("gflops"):
            ref_gpu["speedup_of_this_library"] = round(value / world / ref_gpu["gflops"], 1)
        line.update(extras)
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def ncu_traffic(config, p):
    """dram__bytes_read+write per launch from the committed ncu capture of this config, scaled by the
    batch size (profiles/ncu_traffic.json), or None."""
    f = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        t = json.load(open(f))[config]
        return int(t["dram_bytes_per_item"] * p.nb)
    except Exception:
        return None


def measure_e2e(args, p, tdt, dt, world, rank, local, dist):
    """Same metric through kronmult_batched_host_*: host pointer arrays of host (pinned) vectors."""
    import numpy as np
    import torch
    from kronmult993_b200 import api

    N, d, n, s_el = p.N, p.d, p.n, (8 if dt == "f64" else 4)
    try:
        avail = 0
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable"):
                avail = int(ln.split()[1]) * 1024
        per_item = N * s_el + d * n * n * s_el
        budget = int(avail * 0.45 / max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world))))
        items = min(p.nb, max(1, budget // per_item))
        r = int(p.nb // max(1, p.n_outputs))
        items = max(r, items // r * r)
        n_out = items // r
        h_in = torch.empty(items * N, dtype=tdt).pin_memory()
        h_A = torch.empty(items * d * n * n, dtype=tdt).pin_memory()
        h_out = torch.empty(n_out * N, dtype=tdt).pin_memory()
        h_in.copy_(p.in_slab[: items * N]); h_A.copy_(p.mat_slab[: items * d * n * n]); h_out.copy_(p.out_slab[: n_out * N])
        torch.cuda.synchronize()
        pa = (h_A.data_ptr() + np.arange(items * d, dtype=np.int64) * (n * n * s_el)).astype(np.uint64)
        pi = (h_in.data_ptr() + np.arange(items, dtype=np.int64) * (N * s_el)).astype(np.uint64)
        po = (h_out.data_ptr() + (np.arange(items, dtype=np.int64) // r) * (N * s_el)).astype(np.uint64)

        def call():
            api.kronmult_batched_host(d, n, pa.ctypes.data, n, pi.ctypes.data, po.ctypes.data, 0, items, dtype=tdt,
                                      device=local)
        call()  # warm-up: allocates the staging buffers
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            call()
        torch.cuda.synchronize()
        dt_s = time.perf_counter() - t0
        t = torch.tensor([dt_s], dtype=torch.float64, device=f"cuda:{local}")
        fl = torch.tensor([float(items * 2 * d * n ** (d + 1))], dtype=torch.float64, device=f"cuda:{local}")
        by = torch.tensor([float(items * per_item + items * 8 + n_out * N * s_el), float(n_out * N * s_el)],
                          dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(fl); dist.all_reduce(by)
        sec = float(t.item()) / args.e2e_steps
        return {"value": round(float(fl.item()) / sec * 1e-9, 2), "unit": "GFLOP/s",
                "h2d_bytes_per_step": int(by[0].item()), "d2h_bytes_per_step": int(by[1].item()),
                "ms_per_step": round(sec * 1e3, 2), "steps": args.e2e_steps,
                "h2d_gbs_per_gpu": round(float(by[0].item()) / world / sec * 1e-9, 2),
                "items": int(items * world) if items == p.nb else int(items),
                "note": ("whole per-rank batch" if items == p.nb else f"{items} of {p.nb} items per rank fit host RAM")
                        + "; host pointer arrays -> pinned staging -> H2D -> kernel -> D2H, wall clock, max over ranks"}
    except Exception as ex:
        return {"value": None, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "note": f"e2e failed: {type(ex).__name__}: {ex}"}


if __name__ == "__main__":
    main()
