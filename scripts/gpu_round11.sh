#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_k2n -s 1 -c 1 -f -o gpurun_out/wgrad72_k2n python scripts/profile_conv.py wgrad72 2 > gpurun_out/ncu_wgrad72.log 2>&1
ls -la gpurun_out/wgrad72_k2n.ncu-rep
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 450 -c 200 --csv --log-file gpurun_out/launches7.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench7.log 2>&1
wc -l gpurun_out/launches7.csv
