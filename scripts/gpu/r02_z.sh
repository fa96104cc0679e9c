#!/bin/bash
# round 2, call Z: adversarial fine-tuner on the GPU (first contact)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_adversary_gpu.py -q -m gpu -s 2>&1 | tail -60 | cut -c1-250
