#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/diag_epi.py > gpurun_out/diag_epi.txt 2>&1
head -12 gpurun_out/diag_epi.txt
