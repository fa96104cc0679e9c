#!/bin/bash
# ncu --set full of one convolution launch at the headline size (scripts/profile_conv.py case, kernel regex, output name).
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scripts/gpu/ncu_full.sh wgrad24 wgrad_tc_persistent wgrad24_persistent'
CASE=${1:-wgrad24}; KRE=${2:-wgrad_tc_persistent}; OUT=${3:-$CASE}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 1 -c 1 -f -o gpurun_out/$OUT \
    python scripts/profile_conv.py $CASE 2 > gpurun_out/ncu_$OUT.log 2>&1
ls -la gpurun_out/$OUT.ncu-rep
