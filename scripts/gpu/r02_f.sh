#!/bin/bash
# round 2, call F: whole suite with the hybrid forward + TMA blur, bench A/B of the compensation schemes, per-layer times,
# generator ncu capture with the TMA-staged blur
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader
echo "== TMA-staged blur on its own first (a fault would poison the rest of the process)"
if timeout 600 python -m pytest tests/test_generator_gpu.py tests/test_generator_entry_points_gpu.py -m gpu -q -k "not benchmark_shape" 2>&1 | tail -4 | tee /dev/stderr | grep -q " passed" && ! timeout 600 python -m pytest tests/test_generator_gpu.py -m gpu -q -k "training_defaults" 2>&1 | grep -q "failed\|error"; then
  echo "TMA blur OK"
else
  echo "TMA blur FAILED -> SSR_NO_TMA_BLUR=1 for the rest of this call"; export SSR_NO_TMA_BLUR=1
fi
echo "== gpu tests (all)"
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -25
grep -v "comp \|hybrid " gpurun_out/unet_parity.txt
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== generator only (TMA blur / plain blur)"
timeout 300 python scripts/gen_only.py --size 160 --iters 30
SSR_NO_TMA_BLUR=1 timeout 300 python scripts/gen_only.py --size 160 --iters 30
echo "== bench hybrid / tf32x3"
for sch in hybrid tf32x3; do
  SSR_COMP_SCHEME=$sch timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r02f_bench_$sch.json 2> gpurun_out/r02f_bench_$sch.err
  tail -2 gpurun_out/r02f_bench_$sch.err | cut -c1-300
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02f_bench_$sch.json').read().strip().split('\n')[-1])
print('$sch', d['value'], d['ms_per_step'], 'parity', {k:v for k,v in (d.get('parity') or {}).items() if k not in ('against','bar')}, 'fast', (d.get('fast_mode') or {}).get('value'))
print('   ', {k:(round(v['ms_per_step'],3), round(v['tflops'],1)) for k,v in d['roofline']['per_kind'].items()})
PY
done
SSR_CONV_IMPL=tc3 timeout 300 python scripts/layer_times.py > gpurun_out/r02f_layer_times_hybrid.txt 2>&1; head -18 gpurun_out/r02f_layer_times_hybrid.txt
echo "== generator ncu (TMA blur)"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'blur3d' -s 4 -c 2 -f -o gpurun_out/r02f_blur_tma python scripts/gen_only.py --size 160 --iters 1 --warmup 2 > gpurun_out/r02f_blur_ncu.log 2>&1
ls -la gpurun_out/r02f_blur_tma.ncu-rep
