#!/bin/bash
# knob 13 = ring slots, knob 14 = lag of the persistent n = 8, d = 6 kernel
for tune in "12=0" "12=2,13=4,14=2" "12=2,13=5,14=2" "12=2,13=6,14=2" "12=2,13=6,14=3" "12=2,13=8,14=2" "12=2,13=8,14=3" "12=2,13=8,14=4" "12=1,13=6,14=2"; do
  echo -n "$tune  "
  timeout 60 python tools/fullbench.py --degrees 8 --dims 6 --dtype f64 --target-mb ${MB:-2048} --reps 5 --tune $tune 2>&1 | tail -1 | python -c "
import sys, json
ln = sys.stdin.readline()
try:
    r = json.loads(ln); print(r['path'], r['ms'], r['roofline_frac'])
except Exception:
    print('FAILED', ln[:200])"
done
