#!/bin/bash
# usage: bash tools/ncu_sym5.sh <outname> <config> <scale> <path> [tune]   -- one --set full capture of the n=4,d=5 kernel
out=$1; cfg=$2; scale=$3; path=$4; tune=$5
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sym|wspec' -s 2 -c 1 \
    -f -o gpurun_out/$out python tools/quickbench.py --configs $cfg --scale $scale --reps 1 --path $path ${tune:+--tune $tune} > gpurun_out/ncu_$out.log 2>&1
echo "ncu $out rc=$?"
