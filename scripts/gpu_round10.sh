#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/unet_step_errors.txt
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench.log | cut -c1-330
timeout 300 python scripts/layer_times.py 2>&1 | tail -6 | tee gpurun_out/layer_times3.txt
