#!/bin/bash
mkdir -p gpurun_out
echo "== new test"
timeout 900 python -m pytest tests/test_training_api_gpu.py -m gpu -q -x 2>&1 | tail -15
