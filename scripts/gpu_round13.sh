#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/unet_step_errors.txt
timeout 900 python -m pytest tests/test_unet_gpu.py -m gpu -q -x 2>&1 | tail -3
for s in 1 2 4 0; do
  if [ $s = 0 ]; then unset SSR_WGRAD_SLICES; else export SSR_WGRAD_SLICES=$s; fi
  echo "slices=$s"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c90-200
done
