#!/bin/bash
# round 2, call L (4 GPUs): the driver's scaling line at N = 4
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29621 \
    bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02l_n4.json 2> gpurun_out/r02l_n4.err
tail -4 gpurun_out/r02l_n4.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02l_n4.json').read().strip().split('\n')[-1])
print('n4', d['value'], d['ms_per_step'], d['step_ms'], 'e2e', d['e2e'] and d['e2e']['value'], 'fast', (d.get('fast_mode') or {}).get('value'), 'replicas', d.get('replicas_identical'), 'parity', d.get('parity',{}).get('pred_rel_l2'))
PY
