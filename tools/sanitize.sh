#!/bin/bash
# compute-sanitizer passes over a small slice of the GPU parity suite (memcheck / racecheck / synccheck).
# usage (under gpurun): bash tools/sanitize.sh
mkdir -p gpurun_out
SEL='wspec5_ragged and (63 or 129) or wspec5_factor_layouts or every_kernel_family and (runs or shuffled) and not generic-8 or explicit_plan and 4-5 or edge_batches or unaligned or pairtile_ragged or pairtile_multipass and (7-6 or 10-6) or read_only_shared and (4-3 or 6-4 or 10-5) or pairtile_fp64 and (5-4 or 9-3 or 6-2)'
for tool in memcheck racecheck synccheck; do
  timeout -k 10 1200 compute-sanitizer --tool $tool --error-exitcode 77 --target-processes all \
      python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider -x -k "$SEL" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitize_$tool.log | tail -3
done
