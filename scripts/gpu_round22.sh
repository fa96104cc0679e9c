#!/bin/bash
mkdir -p gpurun_out
echo "== new tests"
timeout 900 python -m pytest tests/test_unet_gpu.py -m gpu -q -x -k "generic_fused or fused_epilogues_equal" 2>&1 | tail -25
echo "== all gpu tests"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "== bench (default)"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
echo "== bench: no generic epilogue fusion"
SSR_NO_EPI_FUSION_GENERIC=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
echo "== layer times"
timeout 300 python scripts/layer_times.py 160 > gpurun_out/layer_times_s6.txt 2>&1; tail -6 gpurun_out/layer_times_s6.txt
