timeout 600 python -m pytest tests/test_unet_gpu.py -m gpu -q -x -k "batched" 2>&1 | tail -5
