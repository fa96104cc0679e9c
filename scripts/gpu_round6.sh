#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_unet_gpu.py -m gpu -q -x 2>&1 | tail -4
timeout 300 python scripts/profile_conv.py all 5 2>&1 | tee gpurun_out/conv_timing_multiissue.txt
timeout 300 python scripts/conv_phase_probe.py fwd24,fwd96 2>&1 | tee gpurun_out/phases_multiissue.txt
