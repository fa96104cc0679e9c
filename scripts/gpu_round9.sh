#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/unet_step_errors.txt
timeout 900 python -m pytest tests/test_unet_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python scripts/layer_times.py 2>&1 | grep -v "^ *[0-9]* wgrad_tc\|wgrad_ref" | tee gpurun_out/layer_times2.txt
