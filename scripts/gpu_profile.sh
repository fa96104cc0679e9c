#!/bin/bash
mkdir -p gpurun_out
python scripts/profile_conv.py all 5 2>&1 | tee gpurun_out/conv_timing.txt
for c in fwd24 wgrad24; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv3d_tc_kernel|wgrad_tc_kernel" -s 1 -c 1 -o gpurun_out/prof_$c -f python scripts/profile_conv.py $c 2 > gpurun_out/ncu_$c.log 2>&1
  tail -2 gpurun_out/ncu_$c.log
done
ls -la gpurun_out/*.ncu-rep
