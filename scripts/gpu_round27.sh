#!/bin/bash
mkdir -p gpurun_out
echo "== gpu tests"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "== bench"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 560 -c 300 --csv --log-file gpurun_out/launches_s8.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-pipeline > gpurun_out/ncu_bench_s8.log 2>&1
wc -l gpurun_out/launches_s8.csv
