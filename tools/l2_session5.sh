#!/bin/bash
# n = 8, d = 5 on the persistent L2-resident kernel: parity, then A/B (knob 12 = 0: pairtile-multipass) and ring / lag (knobs 15 / 16)
timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_abi.py -q -m gpu -x -p no:cacheprovider -k "dmma_l2 or needs_workspace or (read_only and 8-5) or (larger_than_shared and 8-5) or (pairtile_multipass and 8-5)" 2>&1 | tail -3
for tune in ${TUNES:-"12=0" "12=2,15=16,16=10" "12=2,15=16,16=6" "12=2,15=16,16=8" "12=2,15=24,16=12" "12=2,15=24,16=16" "12=2,15=12,16=6" "12=2,15=32,16=20"}; do
  echo -n "$tune  "
  timeout 60 python tools/fullbench.py --degrees 8 --dims 5 --dtype f64 --target-mb ${MB:-2048} --reps 5 --tune $tune 2>&1 | tail -1 | python -c "
import sys, json
ln = sys.stdin.readline()
try:
    r = json.loads(ln); print(r['path'], r['ms'], r['roofline_frac'])
except Exception:
    print('FAILED', ln[:200])"
done
