#!/bin/bash
# GPU session for the pairtile family: parity, then the reference sweep and a throughput sweep.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -p no:cacheprovider -k "pairtile or sweep or golden or named" > gpurun_out/parity_pt.log 2>&1; echo "parity rc=$?"; tail -5 gpurun_out/parity_pt.log
timeout 600 python tools/fullbench.py --levels 9 > gpurun_out/fb_pt_l9.jsonl 2> gpurun_out/fb_pt.err; echo "fb rc=$?"
timeout 900 python tools/fullbench.py --target-mb 2000 --dims 2,3,4,5,6 --max-bytes-item 140000 > gpurun_out/fb_pt_tp.jsonl 2>> gpurun_out/fb_pt.err; echo "fbtp rc=$?"
timeout 600 python tools/fullbench.py --target-mb 2000 --dims 2,3,4,5,6 --max-bytes-item 140000 --dtype f32 > gpurun_out/fb_pt_tp32.jsonl 2>> gpurun_out/fb_pt.err; echo "fbtp32 rc=$?"
tail -3 gpurun_out/fb_pt.err
