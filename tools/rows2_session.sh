#!/bin/bash
# rows2 (kernel_rows2.cuh): parity, then A/B against the kernels it would replace, all (stages, warps) variants
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -p no:cacheprovider -k "rows2" > gpurun_out/rows2_parity.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/rows2_parity.log
T="17=0;17=2,18=0;17=2,18=1;17=2,18=2;17=2,18=3"
timeout 200 python tools/ab_session.py --shapes "5,2;6,2;7,2;8,2;9,2;10,2" --dtypes f64,f32 --tunes "$T" --reset "17=1,18=0" --check > gpurun_out/rows2_ab.jsonl 2> gpurun_out/rows2_ab.err; echo "ab rc=$?"
timeout 100 python tools/ab_session.py --shapes "9,2;10,2" --dtypes f64 --tunes "17=0;17=2,18=0;17=2,18=2" --reset "17=1,18=0" --matrices asgard > gpurun_out/rows2_ab_asgard.jsonl 2>> gpurun_out/rows2_ab.err; echo "ab asgard rc=$?"
timeout 100 python tools/ab_session.py --shapes "9,2;10,2" --dtypes f64 --tunes "17=0;17=2,18=0;17=2,18=2" --reset "17=1,18=0" --matrices reftest > gpurun_out/rows2_ab_reftest.jsonl 2>> gpurun_out/rows2_ab.err; echo "ab reftest rc=$?"
python - <<'PY'
import json
for f in ("rows2_ab", "rows2_ab_asgard", "rows2_ab_reftest"):
    print(f)
    for l in open(f"gpurun_out/{f}.jsonl"):
        r = json.loads(l)
        print(r["dtype"], r["n"], r["d"], " ".join(f"{k}:{v.get('path','?')}/{v.get('frac', v.get('error'))}/{v.get('rel_l2','')}" for k, v in r.items() if isinstance(v, dict)))
PY
tail -5 gpurun_out/rows2_ab.err
