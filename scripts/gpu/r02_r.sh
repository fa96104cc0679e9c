#!/bin/bash
# round 2, call R: deep tiles with the slot wait deferred to the first MMA
mkdir -p gpurun_out
echo "== kernel-level tests"
timeout 600 python -m pytest tests/test_unet_parity_gpu.py tests/test_unet_gpu.py -q -m gpu -x -k "not training_step and not argmax" 2>&1 | tail -4
echo "== layer times new / old tiles"
timeout 300 python scripts/layer_times.py > gpurun_out/r02r_layer_times.txt 2>&1
SSR_TC_OLD_TILES=1 timeout 300 python scripts/layer_times.py > gpurun_out/r02r_layer_times_old.txt 2>&1
paste <(cut -c1-60 gpurun_out/r02r_layer_times.txt) <(cut -c44-60 gpurun_out/r02r_layer_times_old.txt) | head -64
