#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_unet_gpu.py -m gpu -q -x 2>&1 | tail -4
timeout 300 python scripts/profile_conv.py wgrad24,wgrad72,wgrad96 5 2>&1 | tee gpurun_out/conv_timing_k2nall.txt
timeout 300 python scripts/layer_times.py 2>&1 | tee gpurun_out/layer_times.txt
