#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_unet_gpu.py -m gpu -q -x -k "up_parity or fused_epilogues_match or pool_bn or generic_fused or step_parity" > gpurun_out/sanitizer_s2.txt 2>&1
echo "rc=$?"
tail -12 gpurun_out/sanitizer_s2.txt
