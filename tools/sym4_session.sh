#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -p no:cacheprovider -k "half_warp or every_kernel_family or sweep_envelope or random_shapes or unaligned or overlapping or n4d5" 2>&1 | tail -3
timeout 600 python tools/quickbench.py --configs c5_f32 --reps 8 --rounds 2 --fp32-tflops 70.8 --ab "sym5,sym4:10=1,sym4:10=2" 2>/dev/null | cut -c1-200
for path in regtile sym4; do
  for dt in f64 f32; do
    timeout 300 python tools/fullbench.py --degrees 4 --dims 4 --dtype $dt --target-mb 2048 --reps 5 --path $path 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    r = json.loads(ln); print('$path $dt', r['path'], 'ms', r['ms'], 'frac', r['roofline_frac'], 'gbs', r['alg_gbs'])"
  done
done
