#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_unet_gpu.py -m gpu -q -x -k "k2n or wgrad_matches or tc_matches" > gpurun_out/sanitizer_memcheck_full.txt 2>&1
grep -v "Host Frame" gpurun_out/sanitizer_memcheck_full.txt | head -40
