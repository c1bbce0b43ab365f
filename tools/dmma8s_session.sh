#!/bin/bash
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_sharded_gpu.py -q -m gpu -x -p no:cacheprovider -k "sweep_envelope_fp64 and (8-2 or 8-3) or random_shapes or pairtile_fp64 or unaligned or golden or read_only_shared or edge" 2>&1 | tail -3
for tune in 11=0 11=1; do
  timeout 300 python tools/fullbench.py --degrees 8 --dims 2,3 --dtype f64 --target-mb 2000 --reps 5 --tune $tune 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    r = json.loads(ln); print('$tune', r['n'], r['d'], r['path'], 'ms', r['ms'], 'frac', r['roofline_frac'], 'gbs', r['alg_gbs'])"
  timeout 300 python tools/fullbench.py --degrees 8 --dims 2,3 --dtype f64 --levels 9 --reps 5 --tune $tune 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    r = json.loads(ln); print('$tune l9', r['n'], r['d'], r['path'], 'nb', r['nb'], 'ms', r['ms'])"
done
