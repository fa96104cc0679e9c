#!/bin/bash
# N GPUs of one box (gpurun --gpus N): data-parallel bench, then 1 GPU on the same box for the ratio.
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-420
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
