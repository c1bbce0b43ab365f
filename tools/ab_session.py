"""A/B timing of tuning-knob settings on throughput-sized batches, one process, the same problem for every setting.

    python tools/ab_session.py --shapes "9,2;10,2" --tunes "17=0;17=2,18=0;17=2,18=1" [--dtypes f64,f32] [--mb 1000]
                               [--matrices dense|reftest|asgard] [--check]

One JSON line per (dtype, shape): for every setting the kernel family, the best CUDA-event time of the stream-ordered
C-ABI call and the fraction of the applicable roofline (same definition as tools/fullbench.py).  --check compares the
result of every setting on a 2000-item prefix of the batch with the oracle (checker only).  Development tool.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kronmult993_b200 import api, batch  # noqa: E402


def apply(tune, lib):
    for kv in [x for x in tune.split(",") if x]:
        k, v = kv.split("=")
        assert lib.kronmult_b200_set_tuning(int(k), int(v)) == 0, kv


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", required=True)
    ap.add_argument("--tunes", required=True, help="';'-separated settings, each knob=value[,knob=value]")
    ap.add_argument("--reset", default="", help="setting applied after every shape (restores defaults)")
    ap.add_argument("--dtypes", default="f64")
    ap.add_argument("--mb", type=float, default=1000.0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--matrices", default="dense")
    ap.add_argument("--check", action="store_true")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    lib = api.load_library()
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6552.0)
    st = torch.cuda.Stream()
    for name in args.dtypes.split(","):
        dt = torch.float64 if name == "f64" else torch.float32
        peak = (34.1 if name == "f64" else 70.8) * 1e12
        for s in args.shapes.split(";"):
            n, d = (int(x) for x in s.split(","))
            esz = 8 if name == "f64" else 4
            nb = max(64, int(args.mb * 1e6 / (n ** d * esz)))
            p = batch.make_problem(d, n, nb, dt, "cuda", seed=993, alias="runs", items_per_output=32, matrices=args.matrices)
            A, i, o, w = p.pointer_arrays()
            roof = max(p.algorithmic_bytes() / (hbm * 1e9), p.flops() / peak)
            line = {"dtype": name, "n": n, "d": d, "nb": nb, "matrices": args.matrices}
            if args.check:
                from oracle import oracle  # checker only
                nbc = min(nb, 2016)
                hp = batch.make_problem(d, n, nbc, dt, "cpu", seed=5, alias="runs", items_per_output=32,
                                        matrices=args.matrices).to_host()
                exp = oracle.run(hp, "oracle", threads=1)
            for tune in args.tunes.split(";"):
                apply(tune, lib)
                torch.cuda.synchronize()
                best = float("inf")
                try:
                    with torch.cuda.stream(st):
                        for r in range(args.reps + 1):
                            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                            e0.record(st)
                            api.kronmult_batched(d, n, A, p.lda, i, o, w, nb, dtype=dt, stream=st)
                            e1.record(st)
                            e1.synchronize()
                            if r:
                                best = min(best, e0.elapsed_time(e1))
                    res = {"path": api.last_path(), "ms": round(best, 4), "frac": round(roof * 1e3 / best, 3)}
                    if args.check:
                        q = batch.from_host(hp, "cuda")
                        api.run_problem(q)
                        torch.cuda.synchronize()
                        res["rel_l2"] = float(f"{oracle.rel_l2(q.out_slab.cpu().numpy(), exp):.2e}")
                except api.KronmultError as e:
                    res = {"error": str(e)}
                line[tune] = res
                apply(args.reset, lib)
            print(json.dumps(line), flush=True)
            del p, A, i, o, w
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
