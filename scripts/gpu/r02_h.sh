#!/bin/bash
mkdir -p gpurun_out
echo "== compute-sanitizer on the TMA-staged blur"
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python scripts/gen_only.py --size 64 --iters 1 --warmup 0 > gpurun_out/r02h_sanitizer_blur.txt 2>&1
grep -v "^=========     Host Frame\|^=========         in \|^=========                in" gpurun_out/r02h_sanitizer_blur.txt | head -40
echo "== parity file (generator tests use the plain blur in this call)"
SSR_NO_TMA_BLUR=1 timeout 1700 python -m pytest tests/test_unet_parity_gpu.py -m gpu -q -k "96 or 160 or argmax or trained" 2>&1 | tail -15
grep -v "comp \|hybrid " gpurun_out/unet_parity.txt
echo "== ncu: elementwise passes of one step"
SSR_NO_TMA_BLUR=1 timeout 900 ncu --set full --clock-control none -k regex:'head_loss|bn_bwd_apply|pool_bn_bwd|bn_apply|colsum|tf32_split|adam|conv3d_first|wgrad_first' -s 40 -c 30 -f -o gpurun_out/r02h_elementwise \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras --no-pipeline > gpurun_out/r02h_ncu_elementwise.log 2>&1
ls -la gpurun_out/r02h_elementwise.ncu-rep
