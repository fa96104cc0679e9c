#!/bin/bash
# round 2, call B: the compensated forward (conv_impl='tc3'): kernel + step parity at north_star bars, cost vs plain TF32.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader
echo "== parity of the compensated mode"
timeout 1500 python -m pytest tests/test_unet_parity_gpu.py -m gpu -q -x 2>&1 | tail -30
cat gpurun_out/unet_parity.txt
echo "== rest of the gpu suite"
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_unet_parity_gpu.py 2>&1 | tail -12
echo "== seg loss with tc3"
SSR_ENABLE_SEG_LOSS=1 timeout 900 python -m pytest tests/test_seg_loss_gpu.py -m gpu -q 2>&1 | tail -5
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench tc3 / tc"
for impl in tc3 tc; do
  timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --conv-impl $impl > gpurun_out/r02b_bench_$impl.json 2> gpurun_out/r02b_bench_$impl.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r02b_bench_$impl.json'))
print('$impl', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], {k:(round(v['ms_per_step'],3), round(v['tflops'],1)) for k,v in d['roofline']['per_kind'].items()})
PY
  tail -2 gpurun_out/r02b_bench_$impl.err
done
echo "== per-layer times tc3"
SSR_CONV_IMPL=tc3 timeout 300 python scripts/layer_times.py > gpurun_out/r02b_layer_times_tc3.txt 2>&1; tail -8 gpurun_out/r02b_layer_times_tc3.txt
