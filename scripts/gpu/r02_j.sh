#!/bin/bash
# round 2, call J: split-K for the small deep levels -- suite, A/B bench, per-layer times, one clean launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader
echo "== gpu tests"
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -12
echo "== bench A/B: split-K on / off (tc3), fast mode"
for v in "SSR_X=1" "SSR_NO_SPLIT_K=1"; do
  env $v timeout 600 python bench.py --steps 40 --warmup 8 --no-cpu-baseline --no-e2e > gpurun_out/r02j_bench_$v.json 2> gpurun_out/r02j_bench_$v.err
  tail -2 gpurun_out/r02j_bench_$v.err | cut -c1-300
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02j_bench_$v.json').read().strip().split('\n')[-1])
print('$v', round(d['value'],2), round(d['ms_per_step'],3), 'parity', round(d['parity']['pred_rel_l2'],6), round(d['parity']['grad_rel_l2'],5), 'fast', round(d['fast_mode']['value'],2))
print('   ', {k:(round(v['ms_per_step'],3), round(v['tflops'],1)) for k,v in d['roofline']['per_kind'].items()})
PY
done
SSR_CONV_IMPL=tc3 timeout 300 python scripts/layer_times.py > gpurun_out/r02j_layer_times.txt 2>&1; cat gpurun_out/r02j_layer_times.txt | head -60
echo "== launch list (two unpipelined steps)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 600 --csv --log-file gpurun_out/r02j_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-pipeline --no-e2e --no-extras > gpurun_out/r02j_ncu_launches.log 2>&1
wc -l gpurun_out/r02j_launches.csv
