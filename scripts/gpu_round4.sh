#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/unet_step_errors.txt
timeout 300 python -c "
from synthsr_b200._lib import lib
import torch; print('selftest', lib.ssr_tc_selftest(0)); torch.cuda.synchronize()" 2>&1 | tail -2
timeout 600 python -m pytest tests/test_unet_gpu.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 300 python scripts/profile_conv.py all 5 2>&1 | tee gpurun_out/conv_timing.txt
timeout 300 python scripts/conv_phase_probe.py fwd24,fwd96 2>&1 | tee gpurun_out/conv_phases4.txt
timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 ${BENCH_ARGS} 2>&1 | tail -2 | tee gpurun_out/bench.log
