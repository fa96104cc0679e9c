#!/bin/bash
# round 2, call M: bf16x3 compensation (level 5) -- kernel parity, step parity, A/B against the hybrid scheme
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
echo "== kernel tests (levels 4 and 5)"
timeout 900 python -m pytest tests/test_unet_parity_gpu.py -q -m gpu -k "generic_forward or parity_forward" 2>&1 | tail -5
grep -h "ssr_bf16x3_split" gpurun_out/unet_parity.txt | tail -20
echo "== step parity under SSR_COMP_SCHEME=bf16x3"
: > gpurun_out/unet_parity.txt
SSR_COMP_SCHEME=bf16x3 timeout 1500 python -m pytest tests/test_unet_parity_gpu.py tests/test_unet_gpu.py tests/test_predict_gpu.py tests/test_seg_loss_gpu.py -q -m gpu 2>&1 | tail -8
grep -h "^tc3" gpurun_out/unet_parity.txt | tail -12
echo "== A/B"
for sch in hybrid bf16x3; do
  SSR_COMP_SCHEME=$sch timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r02m_bench_$sch.json 2> gpurun_out/r02m_bench_$sch.err
  python - "$sch" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r02m_bench_%s.json'%sys.argv[1]).read().strip().split('\n')[-1])
print(sys.argv[1], d['value'], d['ms_per_step'], d['step_ms'], 'parity', d.get('parity'), 'conv', {k:v for k,v in (d.get('roofline') or {}).items() if k in ('achieved','executed','frac')})
PY
done
SSR_COMP_SCHEME=bf16x3 timeout 300 python scripts/layer_times.py > gpurun_out/r02m_layer_times_bf16x3.txt 2>&1; grep fwd_tc gpurun_out/r02m_layer_times_bf16x3.txt | head -20
