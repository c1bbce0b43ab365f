#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -p no:cacheprovider -k "sweep or tiny or golden or edge or every_kernel or unaligned or overlapping or random_shapes or named or large_batch" 2>&1 | tail -3
for dt in f64 f32; do
for tune in "9=2" "9=1"; do
  timeout 600 python tools/fullbench.py --degrees 2,3,4,5,6,7,8,9,10 --dims 1,2,3,4 --dtype $dt --target-mb 1024 --reps 3 --max-bytes-item 600 --tune $tune 2>/dev/null \
    | python -c "
import sys, json
out=[]
for ln in sys.stdin:
    try: r = json.loads(ln)
    except Exception: continue
    if r['path']=='tiny': out.append('(%d,%d) %.3f' % (r['n'], r['d'], r['roofline_frac']))
print('$dt $tune', ' '.join(out))
"
done
done
