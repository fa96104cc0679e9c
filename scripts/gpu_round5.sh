#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/desc_probe.py 2>&1 | tee gpurun_out/desc_probe.txt
echo "== k2n base"; timeout 200 python scripts/profile_conv.py wgrad24,wgrad72 5 2>&1 | tee gpurun_out/wk_exp.txt
echo "== k2n one dY load (wrong results, bandwidth experiment)"; SSR_WK_EXP=1 timeout 200 python scripts/profile_conv.py wgrad24,wgrad72 5 2>&1 | tee -a gpurun_out/wk_exp.txt
echo "== k2n SA4 SB3"; SSR_WK_CFG=43 timeout 200 python scripts/profile_conv.py wgrad24,wgrad72 5 2>&1 | tee -a gpurun_out/wk_exp.txt
