#!/bin/bash
# sym5 with adjacent lanes sharing a row (knob 5 = 20, development variant): parity on a prefix + A/B at half the C5 batch
mkdir -p gpurun_out
timeout 100 python tools/ab_session.py --shapes "4,5" --dtypes f64,f32 --tunes "5=0;5=20;5=0;5=20" --reset "5=0" --mb 16000 --reps 4 --check > gpurun_out/sym5_pairs_ab.jsonl 2> gpurun_out/sym5_pairs_ab.err; echo "ab rc=$?"
cat gpurun_out/sym5_pairs_ab.jsonl; tail -3 gpurun_out/sym5_pairs_ab.err
