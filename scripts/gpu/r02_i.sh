#!/bin/bash
# round 2, call I: final-state validation + the numbers and ncu evidence quoted in DESIGN.md / BASELINE.md
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader
echo "== gpu tests"
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
echo "== bench (default flags: the line the driver records)"
timeout 900 python bench.py > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; tail -3 gpurun_out/r02i_bench.err | cut -c1-200
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02i_bench.json').read().strip().split('\n')[-1])
print({k: d[k] for k in ('value','ms_per_step','step_ms','gpu_launches','clocks')})
print('e2e', d['e2e']['value'], 'parity', {k:v for k,v in (d.get('parity') or {}).items() if k not in ('against','bar')}, 'fast', (d.get('fast_mode') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
r=d['roofline']; print({k:r[k] for k in ('achieved','frac','frac_of_tf32_peak','all_tc_convolutions','whole_step')})
print({k:(round(v['ms_per_step'],3), round(v['tflops'],1)) for k,v in r['per_kind'].items()})
PY
echo "== other configs"
for c in c1 c4 c5; do
  timeout 600 python bench.py --config $c --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r02i_bench_$c.json 2> gpurun_out/r02i_bench_$c.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02i_bench_$c.json').read().strip().split('\n')[-1])
print('$c', {k: d.get(k) for k in ('value','ms_per_step','gpu_launches')}, 'e2e', (d.get('e2e') or {}).get('value'), 'parity', {k:v for k,v in (d.get('parity') or {}).items() if k not in ('against','bar')}, 'fast', (d.get('fast_mode') or {}).get('value'), 'roof', d['roofline'].get('frac'))
PY
done
echo "== reference arm (3 real 160^3 steps)"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02i_bench_ref.json 2>gpurun_out/r02i_bench_ref.err; cut -c1-200 gpurun_out/r02i_bench_ref.json
echo "== launch list of one unpipelined step"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 400 --csv --log-file gpurun_out/r02i_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-pipeline --no-e2e --no-extras > gpurun_out/r02i_ncu_launches.log 2>&1
wc -l gpurun_out/r02i_launches.csv
echo "== ncu --set full of the step's tensor-core kernels"
for spec in "wgrad:wgrad_tc_persistent:2:2" "generic:conv3d_tc_kernel:20:3" "k2n:conv3d_tc_k2n_kernel:4:4" "upk2n:conv3d_tc_up_k2n_kernel:1:1" "up:conv3d_tc_up_kernel:2:2"; do
  IFS=: read tag kre skip cnt <<< "$spec"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kre -s $skip -c $cnt -f -o gpurun_out/r02i_$tag \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-pipeline --no-e2e --no-extras > gpurun_out/r02i_ncu_$tag.log 2>&1
  ls -la gpurun_out/r02i_$tag.ncu-rep 2>&1 | cut -c20-
done
timeout 600 ncu --set full --clock-control none -k regex:'blur3d' -s 4 -c 2 -f -o gpurun_out/r02i_blur python scripts/gen_only.py --size 160 --iters 1 --warmup 2 > gpurun_out/r02i_ncu_blur.log 2>&1
ls -la gpurun_out/r02i_blur.ncu-rep | cut -c20-
