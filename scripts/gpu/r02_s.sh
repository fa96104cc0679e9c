#!/bin/bash
# round 2, call S: tile choice with the L2 term, ring = two tiles of planes
mkdir -p gpurun_out
echo "== kernel-level tests"
timeout 600 python -m pytest tests/test_unet_parity_gpu.py tests/test_unet_gpu.py -q -m gpu -k "not training_step and not argmax" 2>&1 | tail -4
echo "== layer times new / old tiles"
timeout 300 python scripts/layer_times.py > gpurun_out/r02s_layer_times.txt 2>&1
SSR_TC_OLD_TILES=1 timeout 300 python scripts/layer_times.py > gpurun_out/r02s_layer_times_old.txt 2>&1
paste <(cut -c1-60 gpurun_out/r02s_layer_times.txt) <(cut -c44-60 gpurun_out/r02s_layer_times_old.txt) | head -64
SSR_TC_PRINT_TILES=1 timeout 300 python scripts/layer_times.py 2>&1 | grep "^conv3d_tc" | sort | uniq -c | sort -k2 > gpurun_out/r02s_tiles.txt; cat gpurun_out/r02s_tiles.txt
echo "== A/B"
for v in X OLD_TILES X OLD_TILES; do
  env SSR_TC_$v=1 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r02s_bench_$v.json 2> gpurun_out/r02s_bench_$v.err
  python - "$v" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r02s_bench_%s.json'%sys.argv[1]).read().strip().split('\n')[-1])
print(sys.argv[1], d['value'], d['ms_per_step'], d['step_ms'])
PY
done
