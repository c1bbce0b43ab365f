#!/bin/bash
# verification of the two race fixes of commit f0dd52e: whole GPU suite, then the third sanitizer pass again
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/gpu_all_r02d.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/gpu_all_r02d.log
bash tools/sanitize3.sh
