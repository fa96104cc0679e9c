#!/bin/bash
# A/B of the optional paths inside ONE box (boxes of the pool differ by a few %): default vs each switch off.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/gpu/ab.sh'
run() { echo "== $1"; env $1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c90-190; }
run "SSR_DEFAULT=1"
run "SSR_NO_UP_PARITY=1"
run "SSR_NO_EPI_FUSION=1"
run "SSR_NO_POOL_BN_FUSION=1"
run "SSR_NO_HEAD_BN_SUMS=1"
echo "== --no-pipeline"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-pipeline 2>&1 | tail -1 | cut -c90-190
