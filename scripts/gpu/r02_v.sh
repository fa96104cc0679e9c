#!/bin/bash
# round 2, call V (8 GPUs): the driver's scaling line at N = 8 (and N = 2 on the same box) in the final state
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c
for n in 8 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2963$n \
      bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02v_n$n.json 2> gpurun_out/r02v_n$n.err
  python - $n <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r02v_n%s.json'%n).read().strip().split('\n')[-1])
    print('n'+n, d['value'], d['ms_per_step'], d['step_ms'], 'e2e', d['e2e'] and d['e2e']['value'], 'fast', (d.get('fast_mode') or {}).get('value'), 'replicas', d.get('replicas_identical'), 'parity', d.get('parity',{}).get('pred_rel_l2'))
except Exception as e:
    print('n'+n, 'FAILED', e); print(open('gpurun_out/r02v_n%s.err'%n).read()[-1500:])
PY
done
