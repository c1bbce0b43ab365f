#!/bin/bash
# the library as committed at the end of round 2 (rows2 with even-lane slots): whole GPU suite, then both throughput sweeps again
mkdir -p gpurun_out
timeout 400 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/gpu_all_r02.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/gpu_all_r02.log
timeout 200 python tools/fullbench.py --target-mb 2000 > gpurun_out/fullbench_throughput_f64_r02.jsonl 2> gpurun_out/fullbench_r02.err; echo "fb tp64 rc=$?"
python tools/fbtable.py gpurun_out/fullbench_throughput_f64_r02.jsonl
timeout 200 python tools/fullbench.py --target-mb 2000 --dtype f32 > gpurun_out/fullbench_throughput_f32_r02.jsonl 2>> gpurun_out/fullbench_r02.err; echo "fb tp32 rc=$?"
python tools/fbtable.py gpurun_out/fullbench_throughput_f32_r02.jsonl
