#!/bin/bash
# scratch: ncu --set full of the parity forward kernel and of the first-layer forward kernel
mkdir -p gpurun_out
timeout 300 python scripts/profile_conv.py up72,first 3 2>&1 | tail -6
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3d_tc_up_kernel -s 1 -c 1 -f -o gpurun_out/fwd_up72 python scripts/profile_conv.py up72 1 > gpurun_out/ncu_fwd_up72.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3d_first_kernel -s 1 -c 1 -f -o gpurun_out/first_fwd python scripts/profile_conv.py first 1 > gpurun_out/ncu_first_fwd.log 2>&1
ls -la gpurun_out/*.ncu-rep
