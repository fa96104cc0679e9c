#!/bin/bash
mkdir -p gpurun_out
echo "== unet tests"
timeout 1800 python -m pytest tests/test_unet_gpu.py tests/test_unet_parity_gpu.py tests/test_fullsize_exact_gpu.py tests/test_training_api_gpu.py tests/test_predict_gpu.py -m gpu -q 2>&1 | tail -8
SSR_CONV_IMPL=tc3 timeout 300 python scripts/layer_times.py > gpurun_out/r02k_layer_times.txt 2>&1; grep -n "fwd_tc\|dgrad_tc" gpurun_out/r02k_layer_times.txt | sed -n '6,12p;36,50p'; tail -5 gpurun_out/r02k_layer_times.txt
echo "== bench A/B"
for v in "SSR_X=1" "SSR_NO_SPLIT_K=1"; do
  env $v SSR_BENCH_DUMP_PROF=1 timeout 600 python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-e2e > gpurun_out/r02k_bench_$v.json 2> gpurun_out/r02k_bench_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02k_bench_$v.json').read().strip().split('\n')[-1])
print('$v', round(d['value'],2), round(d['ms_per_step'],3), d['step_ms'], 'parity', round(d['parity']['pred_rel_l2'],6), 'fast', round(d['fast_mode']['value'],2))
print('   ', {k:(round(v['ms_per_step'],3), round(v['tflops'],1)) for k,v in d['roofline']['per_kind'].items()})
PY
  grep "^prof" gpurun_out/r02k_bench_$v.err | awk '$3=="fwd_tc"' | head -34 | awk '{printf "%s ", $4} END {print ""}'
done
