#!/bin/bash
mkdir -p gpurun_out
echo "== new test"
timeout 600 python -m pytest tests/test_unet_gpu.py -m gpu -q -x -k "head_bn_sums or training_step or ref_" 2>&1 | tail -8
echo "== bench"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
echo "== bench no head sums"
SSR_NO_HEAD_BN_SUMS=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
