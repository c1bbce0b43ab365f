#!/bin/bash
# End-of-round measurement session (one GPU): tests, smoke, bench lines, sweeps, launch list, ncu captures, sanitizer.
mkdir -p gpurun_out
{ nvidia-smi; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv; echo; free -g; echo; nproc;
  lscpu | grep -E "Model name|Socket|Core|Thread|^CPU\(s\)"; } > gpurun_out/box.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/gpu_all.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/gpu_all.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
timeout 600 python tools/quickbench.py > gpurun_out/quickbench.jsonl 2> gpurun_out/quickbench.err; echo "quick rc=$?"
timeout 900 python tools/fullbench.py --levels 9 --ref-gpu > gpurun_out/fullbench_l9.jsonl 2> gpurun_out/fullbench.err; echo "fb l9 rc=$?"
timeout 600 python tools/fullbench.py --target-mb 2000 > gpurun_out/fullbench_tp_f64.jsonl 2>> gpurun_out/fullbench.err; echo "fb tp64 rc=$?"
timeout 600 python tools/fullbench.py --target-mb 2000 --dtype f32 > gpurun_out/fullbench_tp_f32.jsonl 2>> gpurun_out/fullbench.err; echo "fb tp32 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1; echo "launch list rc=$?"
bash tools/ncu_shape.sh 4 3 f64 130 prof_c1_tiny
bash tools/ncu_shape.sh 6 5 f64 300 prof_pairtile_n6d5
bash tools/ncu_shape.sh 5 4 f64 300 prof_pairtile_n5d4
bash tools/ncu_shape.sh 10 5 f64 500 prof_pairpass_n10d5
bash tools/sanitize.sh
