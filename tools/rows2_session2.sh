#!/bin/bash
# rows2 after the shifted staging for odd n (fp64) and the aligned-column copy path: parity, sanitizer, A/B
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -p no:cacheprovider -k "rows2" > gpurun_out/rows2_parity2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/rows2_parity2.log
for tool in memcheck racecheck; do
  timeout -k 10 300 compute-sanitizer --tool $tool --error-exitcode 77 --target-processes all \
      python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider -k "rows2 and (float64-9 or float64-10 or float32-7 or float32-8)" > gpurun_out/rows2_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/rows2_$tool.log | tail -3
done
timeout 200 python tools/ab_session.py --shapes "5,2;7,2;9,2" --dtypes f64 --tunes "17=0;17=2,18=0;17=2,18=1" --reset "17=1,18=-1" --check > gpurun_out/rows2_ab2.jsonl 2> gpurun_out/rows2_ab2.err; echo "ab rc=$?"
timeout 100 python tools/ab_session.py --shapes "6,2;8,2;10,2" --dtypes f64 --tunes "17=0;17=2,18=-1" --reset "17=1,18=-1" --matrices asgard --check > gpurun_out/rows2_ab2_asgard.jsonl 2>> gpurun_out/rows2_ab2.err; echo "ab asgard rc=$?"
python - <<'PY'
import json
for f in ("rows2_ab2", "rows2_ab2_asgard"):
    print(f)
    for l in open(f"gpurun_out/{f}.jsonl"):
        r = json.loads(l)
        print(r["dtype"], r["n"], r["d"], " ".join(f"{k}:{v.get('path','?')}/{v.get('frac', v.get('error'))}/{v.get('rel_l2','')}" for k, v in r.items() if isinstance(v, dict)))
PY
tail -5 gpurun_out/rows2_ab2.err
