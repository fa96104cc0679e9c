#!/bin/bash
# scratch: the test subset / A-B of the change being worked on
timeout 900 python -m pytest tests/test_unet_gpu.py -m gpu -q -x -k "plane_linearised or generic_fused or matches_ref" 2>&1 | tail -6
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c90-200
