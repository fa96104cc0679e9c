#!/bin/bash
# round 2, call W: L2-term tile choice on the mid-size (40^3) levels only -- A/B, then the validation of the final state
mkdir -p gpurun_out
for v in X NO_MID_TILES X NO_MID_TILES; do
  env SSR_TC_$v=1 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r02w_bench_$v.json 2> gpurun_out/r02w_bench_$v.err
  python - "$v" <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r02w_bench_%s.json'%sys.argv[1]).read().strip().split('\n')[-1])
print(sys.argv[1], d['value'], d['ms_per_step'], d['step_ms'], {k:round(v['ms_per_step'],3) for k,v in d['roofline']['per_kind'].items()})
PY
done
bash scripts/gpu/validate.sh r02w
