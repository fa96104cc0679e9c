#!/bin/bash
# round 2, call C: whole GPU suite (parity at north_star bars included), the new bench (e2e through training(), live
# parity, fast-mode secondary, percentiles), the other BASELINE configs, both arms.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader
echo "== gpu tests"
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -40
cat gpurun_out/unet_parity.txt
echo "== opt-in suites"
SSR_KERNEL_CROSSCHECK=1 timeout 600 python -m pytest tests/test_generator_entry_points_gpu.py -m gpu -q 2>&1 | tail -3
SSR_ENABLE_SEG_LOSS=1 timeout 900 python -m pytest tests/test_seg_loss_gpu.py -m gpu -q 2>&1 | tail -5
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench (headline, default flags)"
timeout 900 python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; tail -3 gpurun_out/r02c_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02c_bench.json'))
print({k: d[k] for k in ('value','ms_per_step','step_ms','gpu_launches','clocks')})
print('e2e', d['e2e']); print('parity', d.get('parity')); print('fast', d.get('fast_mode')); print('cpu', d.get('cpu_baseline'))
r=d['roofline']; print({k:r[k] for k in ('achieved','frac','frac_of_tf32_peak','all_tc_convolutions','whole_step','kernel')})
print({k:(round(v['ms_per_step'],3), round(v['tflops'],1)) for k,v in r['per_kind'].items()})
PY
echo "== other configs"
for c in c1 c4 c5; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02c_bench_$c.json 2> gpurun_out/r02c_bench_$c.err
  tail -2 gpurun_out/r02c_bench_$c.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r02c_bench_$c.json'))
print('$c', {k: d.get(k) for k in ('value','ms_per_step','step_ms','gpu_launches')}, 'e2e', (d.get('e2e') or {}).get('value'), 'parity', d.get('parity'), 'fast', (d.get('fast_mode') or {}).get('value'), 'roof', d['roofline'].get('frac'))
PY
done
echo "== reference arm"
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02c_bench_ref.json 2>gpurun_out/r02c_bench_ref.err; cut -c1-240 gpurun_out/r02c_bench_ref.json
