#!/bin/bash
# rows2 (kernel_rows2.cuh) development session on one B200: parity of every (T, n) it is built for, racecheck, A/B against the
# kernels it replaces (knob 17 = 0) with 2 and 3 stages (knob 18), dense / ASGarD / lda = 67 layouts, ncu captures of n = 9, 10.
# The variants that were measured and dropped over the round (two-warp CTAs, 1-D TMA per matrix, whole-warp copies, .ca copies)
# are in profiles/rows2_ab_*_r02.jsonl and profiles/rows2_r02.md; their code is in the git history of kernel_rows2.cuh.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -p no:cacheprovider -k "rows2" > gpurun_out/rows2_parity.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/rows2_parity.log
timeout -k 10 300 compute-sanitizer --tool racecheck --error-exitcode 77 --target-processes all \
      python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider -k "rows2 and (9-dt0 or 10-dt0 or 7-dt1 or 8-dt1)" > gpurun_out/rows2_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/rows2_racecheck.log | tail -3
T="17=0;17=2,18=0;17=2,18=1"
timeout 300 python tools/ab_session.py --shapes "5,2;6,2;7,2;8,2;9,2;10,2" --dtypes f64,f32 --tunes "$T" --reset "17=1,18=-1" --check > gpurun_out/rows2_ab.jsonl 2> gpurun_out/rows2_ab.err; echo "ab rc=$?"
for m in asgard reftest; do
  timeout 100 python tools/ab_session.py --shapes "6,2;9,2;10,2" --dtypes f64 --tunes "17=0;17=2,18=-1" --reset "17=1,18=-1" --matrices $m --check > gpurun_out/rows2_ab_$m.jsonl 2>> gpurun_out/rows2_ab.err; echo "ab $m rc=$?"
done
python - <<'PY'
import json
for f in ("rows2_ab", "rows2_ab_asgard", "rows2_ab_reftest"):
    print(f)
    for l in open(f"gpurun_out/{f}.jsonl"):
        r = json.loads(l)
        print(r["dtype"], r["n"], r["d"], " ".join(f"{k}:{v.get('path','?')}/{v.get('frac', v.get('error'))}/{v.get('rel_l2','')}" for k, v in r.items() if isinstance(v, dict)))
PY
if [ -n "$NCU" ]; then
  bash tools/ncu_shape.sh 9 2 f64 500 r02_rows2_n9; bash tools/ncu_shape.sh 10 2 f64 500 r02_rows2_n10
  for f in r02_rows2_n9 r02_rows2_n10; do python tools/ncu_summary.py gpurun_out/$f.ncu-rep > gpurun_out/$f.txt 2>&1; python tools/ncu_smem_ops.py gpurun_out/$f.ncu-rep >> gpurun_out/$f.txt; done
fi
./tools/probes/lds_probe > gpurun_out/lds_probe_r02.jsonl 2>&1
