#!/bin/bash
# One GPU box: what the driver runs at round end -- the full GPU test suite, smoke, the default bench line and the
# reference arm -- plus the launch list of one unpipelined step.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash scripts/gpu/validate.sh [tag]'
TAG=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader
echo "== gpu tests"
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench (default flags)"
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cut -c1-400 gpurun_out/bench_$TAG.json; tail -2 gpurun_out/bench_$TAG.err
echo "== reference arm"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; cut -c1-300 gpurun_out/bench_ref_$TAG.json
echo "== launch list (one unpipelined step under ncu; times are cold-cache and serialised)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-pipeline > gpurun_out/ncu_bench_$TAG.log 2>&1
wc -l gpurun_out/launches_$TAG.csv
