#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/unet_step_errors.txt
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
cat gpurun_out/unet_step_errors.txt
python scripts/profile_conv.py all 5 2>&1 | tee gpurun_out/conv_timing.txt
timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 ${BENCH_ARGS} 2>&1 | tail -3 | tee gpurun_out/bench.log
