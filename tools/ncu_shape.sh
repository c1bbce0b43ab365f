#!/bin/bash
# usage: bash tools/ncu_shape.sh <n> <d> <dtype f64|f32> <target_mb> <outname>
# one --set full capture of the kernel the library picks for one (n, d) (second timed launch)
n=$1; d=$2; dt=$3; mb=$4; out=$5
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'kron_' -s ${NCU_SKIP:-1} -c ${NCU_COUNT:-1} \
    -f -o gpurun_out/$out python tools/fullbench.py --degrees $n --dims $d --dtype $dt --target-mb $mb --reps 1 > gpurun_out/ncu_$out.log 2>&1
echo "ncu $out rc=$?"
