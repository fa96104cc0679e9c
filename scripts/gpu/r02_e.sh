#!/bin/bash
# round 2, call E (2 GPUs): data-parallel exchange -- overlapped prefix all-reduce vs one all-reduce after the backward pass
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
      bench.py --gpus 2 --steps 40 --warmup 8 --no-e2e --no-extras --no-cpu-baseline > gpurun_out/r02e_$tag.json 2> gpurun_out/r02e_$tag.err
  tail -3 gpurun_out/r02e_$tag.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02e_$tag.json').read().strip().split('\n')[-1])
    print('$tag', 'value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],3), 'step_ms', d['step_ms'], 'replicas_identical', d.get('replicas_identical'))
except Exception as e:
    print('$tag failed', e)
PY
}
echo "== new fused-split producers + generator tests (1 GPU)"
timeout 900 python -m pytest tests/test_unet_parity_gpu.py tests/test_generator_gpu.py tests/test_generator_entry_points_gpu.py -m gpu -q -k "producers or 32 or generator or blur or default or training" 2>&1 | tail -6
echo "== 1 GPU reference point (same flags)"
timeout 600 python bench.py --gpus 1 --steps 40 --warmup 8 --no-e2e --no-extras --no-cpu-baseline > gpurun_out/r02e_n1.json 2> gpurun_out/r02e_n1.err
python -c "
import json; d=json.loads(open('gpurun_out/r02e_n1.json').read().strip().split('\n')[-1]); print('n1', round(d['value'],2), round(d['ms_per_step'],3))"
run overlap SSR_DUMMY=1
run single SSR_EXCHANGE_SPLIT_LEVEL=0
run split3 SSR_EXCHANGE_SPLIT_LEVEL=3
run overlap_fast SSR_CONV_IMPL_UNUSED=1
echo "== full default line at 2 GPUs (e2e through training(), fast mode)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 \
    bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r02e_full.json 2> gpurun_out/r02e_full.err
tail -3 gpurun_out/r02e_full.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02e_full.json').read().strip().split('\n')[-1])
print('full', d['value'], d['ms_per_step'], 'e2e', d['e2e'] and d['e2e']['value'], 'fast', (d.get('fast_mode') or {}).get('value'), 'replicas', d.get('replicas_identical'))
PY
