#!/bin/bash
# third compute-sanitizer pass of round 2: the kernels added in the last session (warp-per-item DMMA for n = 5..8, the
# persistent L2-resident kernel for n = 8, d = 5 / 6, the regrouped phases of the n = 8, d = 4 kernel and of the two-kernel route)
mkdir -p gpurun_out
SEL='dmma_warp_per_item and (5-3 or 8-2 or 7-3 or 6-2) or dmma_l2_persistent and (6-9-2 or 6-24-2 or 5-9-2 or 5-40-2) or dmma_l2_launches or reference_large_case or sweep_envelope_fp64 and (8-4 or 7-2) or read_only_shared and (8-5 or 8-6 or 8-4)'
for tool in memcheck racecheck synccheck; do
  timeout -k 10 600 compute-sanitizer --tool $tool --error-exitcode 77 --target-processes all \
      python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider -k "$SEL" > gpurun_out/sanitize3_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/sanitize3_$tool.log | tail -3
done
grep -oE "in kernel_[a-z0-9_]+\.cuh:[0-9]+" gpurun_out/sanitize3_racecheck.log | sort | uniq -c | sort -rn | head
