#!/bin/bash
# One development session on the GPU box: box facts, micro-benchmarks, parity tests, kernel timings.
# Usage (from the repo root, under gpurun): bash tools/gpu_session.sh [steps...]
#   steps: box micro parity quick full smoke   (default: all)
mkdir -p gpurun_out
STEPS="${@:-box micro parity quick full refbin smoke}"
for s in $STEPS; do
  case $s in
    box)
      { nvidia-smi; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv;
        echo; free -g; echo; nproc; lscpu | grep -E "Model name|Socket|Core|Thread|^CPU\(s\)"; } > gpurun_out/box.txt 2>&1 ;;
    micro)
      timeout 300 ./kronmult993_b200/kron_microbench > gpurun_out/microbench.jsonl 2>&1; echo "micro rc=$?" ;;
    parity)
      timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/parity.log 2>&1; echo "parity rc=$?"; tail -5 gpurun_out/parity.log ;;
    quick)
      timeout 600 python tools/quickbench.py > gpurun_out/quickbench.jsonl 2> gpurun_out/quickbench.err; echo "quick rc=$?"; cat gpurun_out/quickbench.jsonl ;;
    full)
      timeout 900 python -m pytest tests/test_fullsize_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/fullsize.log 2>&1; echo "full rc=$?"; tail -5 gpurun_out/fullsize.log ;;
    ncu)
      # one full-section capture of the top kernels (small batch: ncu replays each launch ~40x)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:'regtile4|dmma84|tiny' -s 2 -c 2 \
          -f -o gpurun_out/prof_c3 python tools/quickbench.py --configs c3 --scale 0.05 --reps 1 > gpurun_out/ncu_c3.log 2>&1; echo "ncu c3 rc=$?"
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:'regtile4|dmma84|tiny' -s 2 -c 2 \
          -f -o gpurun_out/prof_c4b python tools/quickbench.py --configs c4b --scale 0.05 --reps 1 > gpurun_out/ncu_c4b.log 2>&1; echo "ncu c4b rc=$?"
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:'regtile4|dmma84|tiny' -s 2 -c 2 \
          -f -o gpurun_out/prof_c5 python tools/quickbench.py --configs c5_f64 --scale 0.02 --reps 1 > gpurun_out/ncu_c5.log 2>&1; echo "ncu c5 rc=$?" ;;
    tune)
      for tv in 0=0 0=1; do echo "tune $tv"; timeout 600 python tools/quickbench.py --configs c3,c5_f64,c5_f32 --tune $tv 2>&1 | tee -a gpurun_out/tune.jsonl; done ;;
    bench)
      timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
      timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json ;;
    refbin)
      timeout 300 python -m pytest tests/test_reference_binaries_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -3
      LD_LIBRARY_PATH=kronmult993_b200 timeout 600 ./oracle/_ref/kronmult_bench_gpu > gpurun_out/ref_bench_gpu_vs_b200.txt 2>&1; echo "ref bench rc=$?"; tail -7 gpurun_out/ref_bench_gpu_vs_b200.txt ;;
    refgpu)
      for c in c5_f64 c3 c4b c2; do timeout 600 python bench.py --impl reference_gpu --config $c --steps 3 --warmup 1 2>> gpurun_out/bench.err | tee -a gpurun_out/bench_refgpu.jsonl; done ;;
    smoke)
      timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -8 gpurun_out/smoke.log ;;
  esac
done
