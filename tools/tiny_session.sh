#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -p no:cacheprovider -k "sweep or tiny or golden or edge or every_kernel or unaligned or overlapping or random_shapes or named" 2>&1 | tail -3
for tune in "9=0" "9=1"; do
  echo "== tune $tune"
  python tools/quickbench.py --configs c1,c2 --reps 10 --tune $tune 2>/dev/null | cut -c1-220
  timeout 600 python tools/fullbench.py --degrees 2,3,4,6,8 --dims 1,2,3,4,5,6 --dtype f64 --target-mb 1024 --reps 3 --max-bytes-item 600 --tune $tune 2>/dev/null \
    | python -c "
import sys, json
for ln in sys.stdin:
    try: r = json.loads(ln)
    except Exception: continue
    if r.get('N',0)*8 >= 128: print(r.get('n'), r.get('d'), r.get('path'), 'ms', r.get('ms'), 'frac', r.get('roofline_frac'))
"
  timeout 600 python tools/fullbench.py --degrees 2,4,8 --dims 2,3,4,5,6 --dtype f32 --target-mb 1024 --reps 3 --max-bytes-item 600 --tune $tune 2>/dev/null \
    | python -c "
import sys, json
for ln in sys.stdin:
    try: r = json.loads(ln)
    except Exception: continue
    if r.get('N',0)*4 >= 128: print('f32', r.get('n'), r.get('d'), r.get('path'), 'ms', r.get('ms'), 'frac', r.get('roofline_frac'))
"
done
