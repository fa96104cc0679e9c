#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-560} -c 300 --csv --log-file gpurun_out/launches_s4.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_s4.log 2>&1
wc -l gpurun_out/launches_s4.csv
