#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/unet_step_errors.txt
timeout 600 python -m pytest tests/test_unet_gpu.py -m gpu -q -x 2>&1 | tail -4
timeout 300 python scripts/layer_times.py 2>&1 | grep "wgrad" | tee gpurun_out/layer_times_wp.txt
