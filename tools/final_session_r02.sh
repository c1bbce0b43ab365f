#!/bin/bash
# End-of-round-2 measurement session (one GPU): smoke, bench lines, sweeps, launch list, ncu captures, sanitizer.
mkdir -p gpurun_out
{ nvidia-smi; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv; echo; free -g; echo; nproc;
  lscpu | grep -E "Model name|Socket|Core|Thread|^CPU\(s\)|NUMA"; } > gpurun_out/box_r02.txt 2>&1
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_r02.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_r02.log
timeout 900 python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_r02.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference_arm_r02.json 2>> gpurun_out/bench_r02.err; cat gpurun_out/bench_reference_arm_r02.json
timeout 900 python tools/fullbench.py --target-mb 2000 > gpurun_out/fullbench_throughput_f64_r02.jsonl 2> gpurun_out/fullbench_r02.err; echo "fb tp64 rc=$?"
timeout 900 python tools/fullbench.py --target-mb 2000 --dtype f32 > gpurun_out/fullbench_throughput_f32_r02.jsonl 2>> gpurun_out/fullbench_r02.err; echo "fb tp32 rc=$?"
timeout 900 python tools/fullbench.py --levels 9 --ref-gpu > gpurun_out/fullbench_l9_r02.jsonl 2>> gpurun_out/fullbench_r02.err; echo "fb l9 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_r02.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_ncu_r02.log 2>&1; echo "launch list rc=$?"
bash tools/ncu_sym5.sh r02_c5f64 c5_f64 0.04 auto
bash tools/ncu_sym5.sh r02_c5f32 c5_f32 0.04 auto
NCU_SKIP=1 bash tools/ncu_one.sh c3 0.05 r02_c3
NCU_SKIP=1 bash tools/ncu_one.sh c1 4 r02_c1
NCU_SKIP=1 bash tools/ncu_one.sh c4b 0.05 r02_c4b
bash tools/sanitize.sh
