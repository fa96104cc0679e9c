#!/bin/bash
# one GPU-box visit: tests, precision probe, bench, ncu launch list.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== precision probe" ; timeout 300 python scripts/tc_precision_probe.py 2>&1 | tail -12 | tee gpurun_out/precision.log
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench" ; timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 2>&1 | tail -8 | tee gpurun_out/bench.log
if [ "${NCU:-1}" = "1" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-480} -c ${NCU_COUNT:-330} --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log; wc -l gpurun_out/launches.csv
fi
