#!/bin/bash
mkdir -p gpurun_out
./tools/probes/lds_probe > gpurun_out/lds_probe_r02.jsonl 2>&1; cat gpurun_out/lds_probe_r02.jsonl
timeout 300 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -p no:cacheprovider -k "rows2" 2>&1 | tail -2
timeout -k 10 300 compute-sanitizer --tool racecheck --error-exitcode 77 --target-processes all \
      python -m pytest tests/test_parity_gpu.py -q -m gpu -p no:cacheprovider -k "rows2 and (9-dt0 or 10-dt0 or 7-dt1 or 8-dt1)" > gpurun_out/rows2_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/rows2_racecheck.log | tail -3
