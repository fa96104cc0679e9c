#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/unet_step_errors.txt
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench.log | cut -c1-400
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-500} -c 200 --csv --log-file gpurun_out/launches6.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench6.log 2>&1
wc -l gpurun_out/launches6.csv
