#!/bin/bash
# round-1 session 2: fused k2n epilogues (BN sums / ELU backward) + fused MaxPool+BN backward: parity, A/B bench, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader
echo "== new tests"
timeout 900 python -m pytest tests/test_unet_gpu.py -m gpu -q -k "fused or pool_bn" 2>&1 | tail -25
echo "== all gpu tests"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15
echo "== bench fused (default)"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_s2_fused.json 2> gpurun_out/bench_s2_fused.err; tail -c 1500 gpurun_out/bench_s2_fused.json
echo "== bench: no epilogue fusion"
SSR_NO_EPI_FUSION=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400
echo "== bench: no pool-bn fusion"
SSR_NO_POOL_BN_FUSION=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400
echo "== bench: neither"
SSR_NO_EPI_FUSION=1 SSR_NO_POOL_BN_FUSION=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400
echo "== layer times"
timeout 300 python scripts/layer_times.py 160 > gpurun_out/layer_times_s2.txt 2>&1; tail -6 gpurun_out/layer_times_s2.txt
echo "== 256^3"
timeout 300 python bench.py --size 256 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-600
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-600} -c 260 --csv --log-file gpurun_out/launches_s2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_s2.log 2>&1
wc -l gpurun_out/launches_s2.csv
