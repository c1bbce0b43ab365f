#!/bin/bash
# usage: bash tools/ncu_one.sh <config> <scale> <outname> [extra quickbench args]
cfg=$1; scale=$2; out=$3; shift 3
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sym|wspec|regtile4|dmma8|tiny|pass_kernel' -s ${NCU_SKIP:-2} -c ${NCU_COUNT:-1} \
    -f -o gpurun_out/$out python tools/quickbench.py --configs $cfg --scale $scale --reps 1 "$@" > gpurun_out/ncu_$out.log 2>&1
echo "ncu $out rc=$?"
