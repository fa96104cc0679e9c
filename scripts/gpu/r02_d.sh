#!/bin/bash
# round 2, call D: hybrid (TF32 + bf16 correction) compensated forward, where the full-size gradient error comes from
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader
echo "== parity file (hybrid scheme default)"
timeout 1700 python -m pytest tests/test_unet_parity_gpu.py -m gpu -q 2>&1 | tail -25
tail -32 gpurun_out/unet_parity.txt
echo "== gradient error diagnostic, 96^3"
timeout 900 python scripts/grad_error_diag.py 96 2>&1 | tail -12
echo "== rest of the suite"
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_unet_parity_gpu.py 2>&1 | tail -8
echo "== bench hybrid / tf32x3 / fast"
for sch in hybrid tf32x3; do
  SSR_COMP_SCHEME=$sch timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r02d_bench_$sch.json 2> gpurun_out/r02d_bench_$sch.err
  tail -2 gpurun_out/r02d_bench_$sch.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02d_bench_$sch.json').read().strip().split('\n')[-1])
print('$sch', d['value'], d['ms_per_step'], 'parity', {k:v for k,v in (d.get('parity') or {}).items() if k not in ('against','bar')}, 'fast', (d.get('fast_mode') or {}).get('value'))
print('   ', {k:(round(v['ms_per_step'],3), round(v['tflops'],1)) for k,v in d['roofline']['per_kind'].items()})
PY
done
SSR_CONV_IMPL=tc3 timeout 300 python scripts/layer_times.py > gpurun_out/r02d_layer_times_hybrid.txt 2>&1; head -18 gpurun_out/r02d_layer_times_hybrid.txt
