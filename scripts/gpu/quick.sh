#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/ku_diag.py > gpurun_out/ku_diag.txt 2>&1
head -60 gpurun_out/ku_diag.txt
