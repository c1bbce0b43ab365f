#!/bin/bash
# rows2: sanitizer passes and ncu captures of n = 9 and n = 10 (fp64)
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout -k 10 300 compute-sanitizer --tool $tool --error-exitcode 77 --target-processes all \
      python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider -k "rows2 and (9-dt0 or 10-dt0 or 7-dt1 or 8-dt1)" > gpurun_out/rows2_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|passed|failed|RACECHECK SUMMARY" gpurun_out/rows2_$tool.log | tail -3
done
bash tools/ncu_shape.sh 9 2 f64 500 r02_rows2_n9
bash tools/ncu_shape.sh 10 2 f64 500 r02_rows2_n10
for f in r02_rows2_n9 r02_rows2_n10; do python tools/ncu_summary.py gpurun_out/$f.ncu-rep > gpurun_out/$f.txt 2>&1; done
head -40 gpurun_out/r02_rows2_n9.txt
