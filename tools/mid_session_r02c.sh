#!/bin/bash
# mid-session check of the whole library: GPU test suite, smoke, fp64 throughput sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/gpu_all_r02c.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gpu_all_r02c.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_r02c.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_r02c.log
timeout 900 python tools/fullbench.py --target-mb 2000 > gpurun_out/fullbench_throughput_f64_r02c.jsonl 2> gpurun_out/fullbench_r02c.err; echo "fb tp64 rc=$?"
python tools/fbtable.py gpurun_out/fullbench_throughput_f64_r02c.jsonl
