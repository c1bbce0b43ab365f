#!/bin/bash
# multi-pass routes: round-1 behaviour (knob 6=0) vs L2-resident chunks (knob 6 MiB) on 1..3 internal streams (knob 8)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -p no:cacheprovider -k "multipass or larger_than_shared or read_only" 2>&1 | tail -3
for tune in ${TUNES:-"6=0" "6=32,7=1,8=3" "6=32,7=0,8=3" "6=16,7=1,8=3" "6=48,7=1,8=2" "6=24,7=1,8=4"}; do
  echo "== tune $tune"
  timeout 600 python tools/fullbench.py --degrees ${DEGS:-6,8,10} --dims ${DIMS:-5,6} --dtype f64 --target-mb 2048 --reps 3 --max-bytes-item 9000000 --tune $tune 2>/dev/null \
    | python -c "
import sys, json
for ln in sys.stdin:
    try: r = json.loads(ln)
    except Exception: continue
    if 'multipass' in r.get('path',''): print(r.get('n'), r.get('d'), r.get('path'), 'ms', r.get('ms'), 'frac', r.get('roofline_frac'))
" | tee -a gpurun_out/multipass_ab.txt
done
