#!/bin/bash
mkdir -p gpurun_out
echo "== activation-gradient diagnostic"
timeout 900 python scripts/actgrad_diag.py 96 noise l2 2>&1 | tail -40 | tee gpurun_out/r02g_actgrad_96_noise_l2.txt
timeout 900 python scripts/actgrad_diag.py 96 gen l1 2>&1 | tail -40 | tee gpurun_out/r02g_actgrad_96_gen_l1.txt
timeout 900 python scripts/actgrad_diag.py 48 noise l2 2>&1 | tail -40 | tee gpurun_out/r02g_actgrad_48_noise_l2.txt
