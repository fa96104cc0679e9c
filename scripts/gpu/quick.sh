#!/bin/bash
# scratch: the test subset / A-B of the change being worked on
echo "== conv tests"
timeout 900 python -m pytest tests/test_unet_gpu.py -m gpu -q -x 2>&1 | tail -8
echo "== bench"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c90-200
echo "== bench SSR_NO_PLANE_TILES"
SSR_NO_PLANE_TILES=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c90-200
timeout 300 python scripts/layer_times.py 160 > gpurun_out/layer_times_s9.txt 2>&1; tail -6 gpurun_out/layer_times_s9.txt
