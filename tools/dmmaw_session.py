"""Development session for the warp-per-item DMMA kernel (kernel_dmma.cuh, kron_dmma8s_kernel): parity of every shape it
supports against the oracle (knob 11 = 2), then A/B timing against the kernels it would replace (knob 11 = 0).

    python tools/dmmaw_session.py [--mb 1000] [--shapes 5,2;5,3;...]
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
from oracle import oracle  # noqa: E402  (checker only)


def parity(n, d, dt):
    worst = 0.0
    for alias, kw, extra in (("runs", dict(items_per_output=5), dict(lda=n + 3)),
                             ("distinct", {}, dict(matrices="reftest")),
                             ("shuffled", dict(items_per_output=4), dict(misalign=1)),
                             ("ref", dict(nb_distinct=1), dict(matrices="asgard"))):
        for nb in (1, 3, 130, 1501):
            hp = batch.make_problem(d, n, nb, dt, "cpu", seed=n * 100 + d * 10 + nb % 7, alias=alias, **kw, **extra).to_host()
            p = batch.from_host(hp, "cuda")
            api.run_problem(p)
            torch.cuda.synchronize()
            assert api.last_path() == "dmma", api.last_path()
            err = oracle.rel_l2(p.out_slab.cpu().numpy(), oracle.run(hp, "oracle", threads=1))
            tol = 1e-12 if dt == torch.float64 else 1e-5
            assert np.isfinite(err) and err <= tol, (n, d, dt, alias, nb, err)
            worst = max(worst, err)
    return worst


def timing(n, d, dt, mb, reps=5):
    esz = 8 if dt == torch.float64 else 4
    nb = max(64, int(mb * 1e6 / (n ** d * esz)))
    p = batch.make_problem(d, n, nb, dt, "cuda", seed=993, alias="runs", items_per_output=32)
    A, i, o, w = p.pointer_arrays()
    st = torch.cuda.Stream()
    torch.cuda.synchronize()  # the problem was built on the default stream; `st` does not wait for it by itself
    best = float("inf")
    with torch.cuda.stream(st):
        for r in range(reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            api.kronmult_batched(d, n, A, p.lda, i, o, w, nb, dtype=dt, stream=st)
            e1.record(st)
            e1.synchronize()
            if r:
                best = min(best, e0.elapsed_time(e1))
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6552.0)
    peak = (34.1 if dt == torch.float64 else 70.8) * 1e12
    roof = max(p.algorithmic_bytes() / (hbm * 1e9), p.flops() / peak)
    return api.last_path(), best, roof * 1e3 / best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=float, default=1000.0)
    ap.add_argument("--shapes", default="5,2;6,2;7,2;8,2;5,3;6,3;7,3;8,3")
    ap.add_argument("--dtypes", default="f64,f32")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--isolate", action="store_true", help="one subprocess per (dtype, shape): a crash costs one line")
    args = ap.parse_args()
    if args.isolate:
        import subprocess
        for name in args.dtypes.split(","):
            for s in args.shapes.split(";"):
                cmd = [sys.executable, __file__, "--mb", str(args.mb), "--shapes", s, "--dtypes", name]
                r = subprocess.run(cmd + (["--no-parity"] if args.no_parity else []), capture_output=True, text=True)
                sys.stdout.write(r.stdout)
                if r.returncode != 0:
                    print(json.dumps({"dtype": name, "shape": s, "failed": r.stderr.strip().splitlines()[-8:]}))
                sys.stdout.flush()
        return
    shapes = [tuple(int(x) for x in s.split(",")) for s in args.shapes.split(";")]
    torch.cuda.set_device(0)
    api.load_library()
    for name in args.dtypes.split(","):
        dt = torch.float64 if name == "f64" else torch.float32
        for n, d in shapes:
            line = {"dtype": name, "n": n, "d": d}
            api.set_tuning(11, 2)
            if not args.no_parity:
                line["rel_l2_worst"] = float(f"{parity(n, d, dt):.3e}")
            pa, ms, fr = timing(n, d, dt, args.mb)
            line.update(new_path=pa, new_ms=round(ms, 4), new_frac=round(fr, 3))
            api.set_tuning(11, 0)
            pa, ms, fr = timing(n, d, dt, args.mb)
            line.update(old_path=pa, old_ms=round(ms, 4), old_frac=round(fr, 3))
            print(json.dumps(line), flush=True)
    api.set_tuning(11, 1)


if __name__ == "__main__":
    main()
