#!/bin/bash
# round 2, call N: bf16x3 as the default scheme; gradient parity with the oracle's pooling routed like the GPU forward
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
: > gpurun_out/unet_parity.txt
echo "== U-Net tests (default scheme)"
timeout 1800 python -m pytest tests/test_unet_parity_gpu.py tests/test_unet_gpu.py tests/test_predict_gpu.py tests/test_seg_loss_gpu.py -q -m gpu 2>&1 | tail -30
grep -h "^tc3\|^ref" gpurun_out/unet_parity.txt | cut -c1-260
echo "== parity tests, hybrid scheme"
: > gpurun_out/unet_parity_hybrid.txt
SSR_COMP_SCHEME=hybrid timeout 1200 python -m pytest tests/test_unet_parity_gpu.py -q -m gpu -k "training_step or argmax" 2>&1 | tail -5
echo "== A/B"
for sch in bf16x3 hybrid; do
  SSR_COMP_SCHEME=$sch timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r02n_bench_$sch.json 2> gpurun_out/r02n_bench_$sch.err
  python - "$sch" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r02n_bench_%s.json'%sys.argv[1]).read().strip().split('\n')[-1])
print(sys.argv[1], d['value'], d['ms_per_step'], d['step_ms'], 'parity', d.get('parity',{}).get('pred_rel_l2'), 'conv', {k:v for k,v in (d.get('roofline') or {}).items() if k in ('achieved','executed','frac')})
PY
done
timeout 300 python scripts/layer_times.py > gpurun_out/r02n_layer_times.txt 2>&1; grep "fwd_tc" gpurun_out/r02n_layer_times.txt | head -20
