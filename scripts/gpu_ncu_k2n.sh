#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3d_tc_k2n -s 1 -c 1 -f -o gpurun_out/fwd24_k2n python scripts/profile_conv.py fwd24 2 > gpurun_out/ncu_fwd24_k2n.log 2>&1
ls -la gpurun_out/fwd24_k2n.ncu-rep
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-330
