#!/bin/bash
# n = 8, d = 6 persistent L2-resident kernel: parity, then A/B of knob 12 (0 = two-kernel route, 1 = fused, 2 = fused + L2 hint),
# knob 13 = ring slots, knob 14 = lag
timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_abi.py -q -m gpu -x -p no:cacheprovider -k "dmma_l2 or large_case or needs_workspace or (read_only and 8-6)" 2>&1 | tail -3
for tune in ${TUNES:-"12=0" "12=2,13=4,14=2" "12=2,13=5,14=2" "12=2,13=5,14=3" "12=2,13=6,14=2" "12=2,13=6,14=3" "12=2,13=6,14=4" "12=2,13=8,14=3" "12=2,13=8,14=4" "12=2,13=8,14=5"}; do
  echo -n "$tune  "
  timeout 60 python tools/fullbench.py --degrees 8 --dims 6 --dtype f64 --target-mb ${MB:-2048} --reps 5 --tune $tune 2>&1 | tail -1 | python -c "
import sys, json
ln = sys.stdin.readline()
try:
    r = json.loads(ln); print(r['path'], r['ms'], r['roofline_frac'])
except Exception:
    print('FAILED', ln[:200])"
done
