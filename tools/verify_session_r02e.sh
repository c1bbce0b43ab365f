#!/bin/bash
# whole GPU suite again (the first run of the final session stopped at a stale path assertion), then the .ca copy A/B of rows2
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/gpu_all_r02.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gpu_all_r02.log
timeout 150 python tools/ab_session.py --shapes "6,2;9,2;10,2" --dtypes f64 --tunes "17=2,18=0;17=2,18=1;17=2,18=2;17=2,18=3" --reset "17=1,18=-1" --check > gpurun_out/rows2_ab7.jsonl 2> gpurun_out/rows2_ab7.err; echo "ab rc=$?"
timeout 100 python tools/ab_session.py --shapes "8,2;10,2" --dtypes f32 --tunes "17=2,18=1;17=2,18=3" --reset "17=1,18=-1" --check >> gpurun_out/rows2_ab7.jsonl 2>> gpurun_out/rows2_ab7.err; echo "ab rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/rows2_ab7.jsonl"):
    r = json.loads(l)
    print(r["dtype"], r["n"], r["d"], " ".join(f"{k}:{v.get('path','?')}/{v.get('frac', v.get('error'))}/{v.get('rel_l2','')}" for k, v in r.items() if isinstance(v, dict)))
PY
tail -3 gpurun_out/rows2_ab7.err
