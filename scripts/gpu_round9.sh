#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/unet_step_errors.txt
timeout 300 python -m pytest tests/test_unet_gpu.py -m gpu -q -x -k "k2n_forward" 2>&1 | tail -15
timeout 300 python scripts/profile_conv.py fwd24 5 2>&1 | tail -3
SSR_NO_FWD_K2N=1 timeout 300 python scripts/profile_conv.py fwd24 5 2>&1 | tail -3
