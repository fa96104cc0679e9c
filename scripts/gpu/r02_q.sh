#!/bin/bash
# round 2, call Q: ring of plane accumulators + deep tiles in the generic convolution kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
echo "== kernel-level tests"
timeout 600 python -m pytest tests/test_unet_parity_gpu.py tests/test_unet_gpu.py -q -m gpu -x -k "not training_step and not argmax" 2>&1 | tail -8
echo "== tile choices (one step)"
SSR_TC_PRINT_TILES=1 timeout 300 python scripts/layer_times.py 2>&1 | grep "^conv3d_tc" | sort | uniq -c | sort -k2 > gpurun_out/r02q_tiles.txt; cat gpurun_out/r02q_tiles.txt
echo "== layer times new / old tiles"
timeout 300 python scripts/layer_times.py > gpurun_out/r02q_layer_times.txt 2>&1
SSR_TC_OLD_TILES=1 timeout 300 python scripts/layer_times.py > gpurun_out/r02q_layer_times_old.txt 2>&1
paste <(cut -c1-60 gpurun_out/r02q_layer_times.txt) <(cut -c44-60 gpurun_out/r02q_layer_times_old.txt) | head -64
echo "== A/B"
for v in X OLD_TILES; do
  env SSR_TC_$v=1 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r02q_bench_$v.json 2> gpurun_out/r02q_bench_$v.err
  python - "$v" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r02q_bench_%s.json'%sys.argv[1]).read().strip().split('\n')[-1])
print(sys.argv[1], d['value'], d['ms_per_step'], d['step_ms'], 'parity', d.get('parity',{}).get('pred_rel_l2'))
PY
done
