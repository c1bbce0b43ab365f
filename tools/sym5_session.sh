#!/bin/bash
# A/B session for the symmetric n=4, d=5 kernel (library built with -DKRON_SYM5_VARIANTS)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -p no:cacheprovider -k "n4d5 or every_kernel_family" > gpurun_out/parity_sym5.log 2>&1; echo "parity rc=$?"; tail -4 gpurun_out/parity_sym5.log
timeout 900 python tools/quickbench.py --configs c5_f64 --reps 8 --rounds 2 --fp64-tflops 34.1 \
   --ab "${AB64:-wspec5,sym5:5=0,sym5:5=4,sym5:5=6,sym5:5=7,sym5:5=9}" > gpurun_out/ab_sym5_f64.jsonl 2> gpurun_out/ab_sym5_f64.err; echo "ab f64 rc=$?"; cut -c1-200 gpurun_out/ab_sym5_f64.jsonl; tail -3 gpurun_out/ab_sym5_f64.err
timeout 600 python tools/quickbench.py --configs c5_f32 --reps 8 --rounds 2 --fp32-tflops 70.8 \
   --ab "${AB32:-wspec5,sym5:5=0,sym5:5=2,sym5:5=6,sym5:5=7,sym5:5=8}" > gpurun_out/ab_sym5_f32.jsonl 2> gpurun_out/ab_sym5_f32.err; echo "ab f32 rc=$?"; cut -c1-200 gpurun_out/ab_sym5_f32.jsonl; tail -3 gpurun_out/ab_sym5_f32.err
