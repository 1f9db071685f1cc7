#!/bin/bash
# ncu --set full of one case of the conv harness: tools/ncu_conv_case.sh <case index> <kernel regex> <out prefix>
# (run under gpurun; the first matching launch is the eager verification launch of that case)
set -e
idx=$1; re=$2; out=$3
ncu --set full --clock-control none --import-source on -k regex:$re -c 1 -o gpurun_out/$out -f \
    ./tests/cuda/test_conv_exe.so big $idx > gpurun_out/$out.log 2>&1
ncu -i gpurun_out/$out.ncu-rep --page raw --csv > gpurun_out/$out.raw.csv
