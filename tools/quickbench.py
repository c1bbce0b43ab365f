"""Kernel-only timing of every BASELINE.json configuration on one GPU (development tool).

Prints one JSON line per configuration: CUDA-event time of the stream-ordered C-ABI call, GFLOP/s,
algorithmic GB/s and the fraction of the applicable roofline (HBM from MEASURED_PEAKS.json, FP peak
from --fp64-tflops / --fp32-tflops).  bench.py is the graded benchmark; this is for iteration.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kronmult993_b200 import api, batch  # noqa: E402

CONFIGS = {
    "c1": (3, 4, 65536, torch.float64, 1),
    "c2": (2, 2, 1 << 24, torch.float64, 1),
    "c3": (6, 4, 1 << 20, torch.float64, 32),
    "c4a": (4, 8, 1 << 19, torch.float64, 1),
    "c4b": (4, 8, 1 << 19, torch.float64, 32),
    "c5_f64": (5, 4, 1 << 23, torch.float64, 32),
    "c5_f32": (5, 4, 1 << 23, torch.float32, 32),
    # the reference's named cases (tests/kronmult_bench_gpu.cpp:68-72): stride 67, 5 distinct outputs
    "ref_medium": (3, 6, 384, torch.float64, "ref"),
    "ref_large": (6, 8, 896, torch.float64, "ref"),
    "ref_realistic": (6, 8, 3903, torch.float64, "ref"),
    "ref_realistic_f32": (6, 8, 3903, torch.float32, "ref"),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c2,c3,c4a,c4b,c5_f64,c5_f32,c1")
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the batch")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--path", default="auto")
    ap.add_argument("--fp64-tflops", type=float, default=37.0)
    ap.add_argument("--fp32-tflops", type=float, default=75.0)
    ap.add_argument("--alias", default=None)
    ap.add_argument("--plan", action="store_true", help="build an aliasing plan first and time its execution")
    ap.add_argument("--telemetry", action="store_true", help="NVML SM clock / power after every rep")
    ap.add_argument("--tune", default=None, help="knob=value[,knob=value] for kronmult_b200_set_tuning")
    ap.add_argument("--ab", default=None,
                    help="A/B list 'path[:knob=value[:knob=value]],...' timed on the SAME problem, interleaved --rounds times")
    ap.add_argument("--rounds", type=int, default=1)
    ap.add_argument("--share-inputs", action="store_true",
                    help="ASGarD-style shared inputs through kronmult_batched_const: item t of output group i reads "
                         "input vector (i + t) mod #outputs, so every vector is read by r items of r consecutive groups")
    args = ap.parse_args()
    hbm = 6552.3
    mp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(mp):
        hbm = json.load(open(mp)).get("hbm_gbs", hbm)
    torch.cuda.set_device(0)
    nv = None
    if args.telemetry:
        import pynvml
        pynvml.nvmlInit()
        nv = pynvml.nvmlDeviceGetHandleByIndex(0)
    stream = torch.cuda.Stream()
    if args.tune:
        for kv in args.tune.split(","):
            kn, va = kv.split("=")
            api.set_tuning(int(kn), int(va))
    for name in args.configs.split(","):
        d, n, nb, dt, r = CONFIGS[name]
        nb = max(1, int(nb * args.scale))
        if r == "ref":
            p = batch.make_problem(d, n, nb, dt, "cuda", seed=993, alias="ref", nb_distinct=5, matrices="reftest")
        else:
            alias = args.alias or ("runs" if r > 1 else "distinct")
            p = batch.make_problem(d, n, nb, dt, "cuda", seed=993, alias=alias, items_per_output=r)
        n_in = p.nb
        if args.share_inputs and r != "ref" and r > 1:
            k = torch.arange(p.nb, device="cuda")
            n_in = p.n_outputs
            p.in_off = ((k // r + k % r) % n_in) * p.N
        A, i, o, w = p.pointer_arrays()
        if args.share_inputs and api.needs_workspace(p.d, p.n, dt):
            p.alloc_workspaces()
            A, i, o, w = p.pointer_arrays()
        api.force_path(args.path)
        torch.cuda.synchronize()  # the problem was built on the default stream
        if args.ab:
            fl, by = p.flops(), p.algorithmic_bytes()
            peak = (args.fp64_tflops if dt == torch.float64 else args.fp32_tflops) * 1e12
            roof = max(by / (hbm * 1e9), fl / peak)
            for rnd in range(args.rounds):
                for ent in args.ab.split(","):
                    parts = ent.split(":")
                    api.force_path(parts[0])
                    for kv in parts[1:]:
                        kn, va = kv.split("=")
                        api.set_tuning(int(kn), int(va))
                    ts = []
                    with torch.cuda.stream(stream):
                        for rep in range(args.reps + 1):
                            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                            e0.record(stream)
                            api.kronmult_batched(p.d, p.n, A, p.lda, i, o, w, p.nb, dtype=dt, stream=stream)
                            e1.record(stream)
                            e1.synchronize()
                            if rep >= 1:
                                ts.append(e0.elapsed_time(e1))
                    ts_s = sorted(ts)
                    print(json.dumps({"config": name, "ab": ent, "round": rnd, "path": api.last_path(), "nb": nb,
                                      "ms_min": round(ts_s[0], 4), "ms_med": round(ts_s[len(ts_s) // 2], 4),
                                      "ms_last": round(ts[-1], 4), "frac_min": round(roof * 1e3 / ts_s[0], 4),
                                      "frac_med": round(roof * 1e3 / ts_s[len(ts_s) // 2], 4)}), flush=True)
            api.force_path("auto")
            del p, A, i, o, w
            torch.cuda.empty_cache()
            continue
        times, tele = [], []
        plan, plan_info = None, {}
        if args.plan:
            import time
            t0 = time.perf_counter()
            plan = api.Plan(p.d, p.n, A, p.lda, i, o, p.nb, dtype=dt)
            plan_info = {"plan_build_ms": round((time.perf_counter() - t0) * 1e3, 3), **plan.stats()}
        with torch.cuda.stream(stream):
            for rep in range(args.reps + 2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                if plan is not None:
                    plan.execute(stream)
                elif args.share_inputs:
                    api.kronmult_batched_const(p.d, p.n, A, p.lda, i, o, w, p.nb, dtype=dt, stream=stream)
                else:
                    api.kronmult_batched(p.d, p.n, A, p.lda, i, o, w, p.nb, dtype=dt, stream=stream)
                e1.record(stream)
                e1.synchronize()
                if rep >= 2:
                    times.append(e0.elapsed_time(e1))
                if nv is not None:
                    tele.append((pynvml.nvmlDeviceGetClockInfo(nv, pynvml.NVML_CLOCK_SM),
                                 round(pynvml.nvmlDeviceGetPowerUsage(nv) / 1000)))
        api.force_path("auto")
        t = min(times) * 1e-3
        fl, by = p.flops(), p.algorithmic_bytes()
        if n_in != p.nb:  # every distinct input vector is compulsory traffic once
            by -= (p.nb - n_in) * p.N * (8 if dt == torch.float64 else 4)
        peak = (args.fp64_tflops if dt == torch.float64 else args.fp32_tflops) * 1e12
        roof = max(by / (hbm * 1e9), fl / peak)
        print(json.dumps({"config": name + ("+shared_inputs" if n_in != p.nb else ""), "path": api.last_path(), "nb": nb, "ms": round(t * 1e3, 4),
                          "ms_all": [round(x, 4) for x in times], "gflops": round(fl / t * 1e-9, 1),
                          "alg_gbs": round(by / t * 1e-9, 1), "roofline_ms": round(roof * 1e3, 4),
                          "frac": round(roof / t, 4), "bound": "hbm" if by / (hbm * 1e9) >= fl / peak else "fp",
                          **plan_info, **({"sm_mhz_power_w": tele} if tele else {})}),
              flush=True)
        del p, A, i, o, w
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
