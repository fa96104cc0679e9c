#!/bin/bash
# One GPU box: full GPU test suite, smoke, headline bench (with the CPU baseline), launch list of one step.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/gpu/validate.sh [tag]'
TAG=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader
echo "== gpu tests"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "== generator entry points vs the host emulator (opt-in cross-check)"
SSR_KERNEL_CROSSCHECK=1 timeout 600 python -m pytest tests/test_generator_entry_points_gpu.py -m gpu -q 2>&1 | tail -8
echo "== segmentation-regularised loss (opt-in, first GPU contact)"
SSR_ENABLE_SEG_LOSS=1 timeout 900 python -m pytest tests/test_seg_loss_gpu.py -m gpu -q 2>&1 | tail -12
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cut -c1-330 gpurun_out/bench_$TAG.json; tail -2 gpurun_out/bench_$TAG.err
echo "== reference arm"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>&1; cut -c1-300 gpurun_out/bench_ref_$TAG.json
echo "== launch list (one unpipelined step under ncu; times are cold-cache and serialised)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 520 -c 300 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-pipeline > gpurun_out/ncu_bench_$TAG.log 2>&1
wc -l gpurun_out/launches_$TAG.csv
