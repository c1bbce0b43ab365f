#!/bin/bash
# rows2 with slots padded to an even number of lanes (odd n): parity, racecheck, A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -p no:cacheprovider -k "rows2 or (sweep_envelope and (9-2 or 7-2 or 5-2))" 2>&1 | tail -2
timeout 150 python tools/ab_session.py --shapes "5,2;7,2;9,2" --dtypes f64 --tunes "17=0;17=2,18=0;17=2,18=1" --reset "17=1,18=-1" --check > gpurun_out/rows2_ab8.jsonl 2> gpurun_out/rows2_ab8.err; echo "ab rc=$?"
timeout 100 python tools/ab_session.py --shapes "7,2;9,2" --dtypes f32 --tunes "17=0;17=2,18=0" --reset "17=1,18=-1" --check >> gpurun_out/rows2_ab8.jsonl 2>> gpurun_out/rows2_ab8.err; echo "ab rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/rows2_ab8.jsonl"):
    r = json.loads(l)
    print(r["dtype"], r["n"], r["d"], " ".join(f"{k}:{v.get('path','?')}/{v.get('frac', v.get('error'))}/{v.get('rel_l2','')}" for k, v in r.items() if isinstance(v, dict)))
PY
tail -3 gpurun_out/rows2_ab8.err
timeout -k 10 200 compute-sanitizer --tool racecheck --error-exitcode 77 --target-processes all \
      python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider -k "rows2 and (9-dt0 or 7-dt1)" > gpurun_out/rows2_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/rows2_racecheck.log | tail -3
