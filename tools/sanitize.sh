#!/bin/bash
# compute-sanitizer passes over a small slice of the GPU parity suite (memcheck / racecheck / synccheck).
# usage (under gpurun): bash tools/sanitize.sh
mkdir -p gpurun_out
SEL='wspec5_ragged and (63 or 129) or wspec5_factor_layouts or every_kernel_family and (runs or shuffled) and not generic-8 or explicit_plan and 4-5 or edge_batches or unaligned'
for tool in memcheck racecheck synccheck; do
  timeout -k 10 1200 compute-sanitizer --tool $tool --error-exitcode 77 --target-processes all \
      python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider -x -k "$SEL" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitize_$tool.log | tail -3
done
