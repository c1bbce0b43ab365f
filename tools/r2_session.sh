#!/bin/bash
# round-2 GPU session: new tests (host path, sharded single-rank, ASGarD builder, reference programs)
mkdir -p gpurun_out
( time timeout 1700 ./oracle/_ref/kronmult_fullbench_gpu ) > gpurun_out/ref_fullbench_gpu.txt 2>&1 &
FB=$!
KRON_SKIP_FULLBENCH=1 timeout 900 python -m pytest tests/test_sharded_gpu.py tests/test_reference_binaries_gpu.py tests/test_multigpu_nccl.py -q -m gpu -x -p no:cacheprovider -rs > gpurun_out/r2_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2_tests.log
wait $FB; echo "fullbench rc=$?"; tail -4 gpurun_out/ref_fullbench_gpu.txt
