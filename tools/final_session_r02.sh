#!/bin/bash
# End-of-round-2 measurements with the final library (one GPU): whole GPU suite, smoke, bench line (N = 1), reference arm,
# throughput sweeps, launch list of the bench, shared-memory probe.  (ncu captures: tools/ncu_sym5.sh, tools/ncu_one.sh,
# tools/ncu_shape.sh; sanitizer: tools/sanitize.sh, tools/sanitize3.sh; rows2: tools/rows2_session.sh.)
mkdir -p gpurun_out
{ nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv; nproc; } > gpurun_out/box_r02c.txt 2>&1
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/gpu_all_r02.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/gpu_all_r02.log
timeout 200 python __graft_entry__.py --smoke > gpurun_out/smoke_r02.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_r02.log
timeout 600 python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_r02.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference_arm_r02.json 2>> gpurun_out/bench_r02.err; echo "ref arm rc=$?"
timeout 400 python tools/fullbench.py --target-mb 2000 > gpurun_out/fullbench_throughput_f64_r02.jsonl 2> gpurun_out/fullbench_r02.err; echo "fb tp64 rc=$?"
python tools/fbtable.py gpurun_out/fullbench_throughput_f64_r02.jsonl
timeout 400 python tools/fullbench.py --target-mb 2000 --dtype f32 > gpurun_out/fullbench_throughput_f32_r02.jsonl 2>> gpurun_out/fullbench_r02.err; echo "fb tp32 rc=$?"
python tools/fbtable.py gpurun_out/fullbench_throughput_f32_r02.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench_r02.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_ncu_r02.log 2>&1; echo "launch list rc=$?"
./tools/probes/lds_probe > gpurun_out/lds_probe_r02.jsonl 2>&1; echo "probe rc=$?"
