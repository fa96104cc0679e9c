#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/unet_step_errors.txt
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
tail -4 gpurun_out/unet_step_errors.txt
echo "== bench N=1"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_n1.log
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 420 -c 230 --csv --log-file gpurun_out/launches2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench2.log 2>&1
wc -l gpurun_out/launches2.csv
