#!/bin/bash
mkdir -p gpurun_out
echo "== predict tests"
timeout 900 python -m pytest tests/test_predict_gpu.py -m gpu -q -x 2>&1 | tail -15
cat gpurun_out/real_weights_inference.txt
echo "== CLI"
timeout 600 python scripts/predict_command_line.py baseline/_ref/images/brain1.nii.gz gpurun_out/brain1_SynthSR.nii.gz --model baseline/_ref/models/SynthSR_v10_210712.h5 2>&1 | tail -5
ls -la gpurun_out/brain1_SynthSR.nii.gz
