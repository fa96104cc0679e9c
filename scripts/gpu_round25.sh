#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
echo "== 2 GPUs"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-420
echo "== 1 GPU"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
echo "== c4"
timeout 300 python scripts/c4_probe.py 2>&1 | tail -3
