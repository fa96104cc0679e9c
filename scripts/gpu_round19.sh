#!/bin/bash
# k2n accumulate prefetch + 3x3x3 blur kernel: tests, bench, launch list; ncu --set full of the dominant kernel (wgrad 24->24 @160^3)
mkdir -p gpurun_out
echo "== gpu tests"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "== bench"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 560 -c 300 --csv --log-file gpurun_out/launches_s5.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_s5.log 2>&1
wc -l gpurun_out/launches_s5.csv
echo "== ncu full: wgrad 24->24"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_persistent -s 1 -c 1 -f -o gpurun_out/wgrad24_persistent python scripts/profile_conv.py wgrad24 2 > gpurun_out/ncu_wgrad24.log 2>&1
ls -la gpurun_out/wgrad24_persistent.ncu-rep
