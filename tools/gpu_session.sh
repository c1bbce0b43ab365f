#!/bin/bash
# One development session on the GPU box: box facts, micro-benchmarks, parity tests, kernel timings.
# Usage (from the repo root, under gpurun): bash tools/gpu_session.sh [steps...]
#   steps: box micro parity quick full smoke   (default: all)
mkdir -p gpurun_out
STEPS="${@:-box micro parity quick full smoke}"
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
    smoke)
      timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -8 gpurun_out/smoke.log ;;
  esac
done
