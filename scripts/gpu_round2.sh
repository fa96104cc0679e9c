#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/unet_step_errors.txt
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
cat gpurun_out/unet_step_errors.txt | tail -4
echo "== conv timing"; python scripts/profile_conv.py all 5 2>&1 | tee gpurun_out/conv_timing.txt
echo "== bench N=1"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_n1.log
if [ "${NGPU:-1}" -ge 2 ]; then
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_n2.log
fi
