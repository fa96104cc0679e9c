#!/bin/bash
# parity-decomposed decoder convolutions (fwd + dgrad): kernel parity, step parity, full gpu suite, A/B bench
mkdir -p gpurun_out
echo "== new tests"
timeout 900 python -m pytest tests/test_unet_gpu.py -m gpu -q -x -k "parity or up_" 2>&1 | tail -25
echo "== all gpu tests"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15
echo "== bench (default: parity on)"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
echo "== bench: parity wgrad off"
SSR_NO_UP_WGRAD=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
echo "== layer times"
timeout 300 python scripts/layer_times.py 160 > gpurun_out/layer_times_s3.txt 2>&1; tail -6 gpurun_out/layer_times_s3.txt
