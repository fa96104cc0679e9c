#!/bin/bash
# round 2, call X: the other BASELINE configs in the final state (bf16x3)
mkdir -p gpurun_out
for c in c4 c5 c1; do
  timeout 900 python bench.py --config $c --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02x_bench_$c.json 2> gpurun_out/r02x_bench_$c.err
  python - $c <<'PY'
import json,sys
c=sys.argv[1]
d=json.loads(open('gpurun_out/r02x_bench_%s.json'%c).read().strip().split('\n')[-1])
print(c, {k: d.get(k) for k in ('value','ms_per_step','gpu_launches')}, 'e2e', (d.get('e2e') or {}).get('value'), 'parity', {k:v for k,v in (d.get('parity') or {}).items() if k not in ('against','bar')}, 'fast', (d.get('fast_mode') or {}).get('value'), 'roof', (d.get('roofline') or {}).get('frac'))
PY
done
