#!/bin/bash
# last measurements of round 2 with the final library: tests, smoke, bench (N=1), throughput sweeps, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/gpu_all_r02.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/gpu_all_r02.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_r02.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_r02.log
timeout 900 python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; echo "bench rc=$?"
timeout 900 python tools/fullbench.py --target-mb 2000 > gpurun_out/fullbench_throughput_f64_r02.jsonl 2> gpurun_out/fullbench_r02.err; echo "fb tp64 rc=$?"
timeout 900 python tools/fullbench.py --target-mb 2000 --dtype f32 > gpurun_out/fullbench_throughput_f32_r02.jsonl 2>> gpurun_out/fullbench_r02.err; echo "fb tp32 rc=$?"
timeout 900 python tools/fullbench.py --levels 9 --ref-gpu > gpurun_out/fullbench_l9_r02.jsonl 2>> gpurun_out/fullbench_r02.err; echo "fb l9 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench_r02.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_ncu_r02.log 2>&1; echo "launch list rc=$?"
