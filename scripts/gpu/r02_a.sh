#!/bin/bash
# round 2, call A: state of HEAD on a B200 (suite, opt-in suites), generator ncu evidence, the other BASELINE configs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader
nproc; free -g | head -2
echo "== gpu tests"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
echo "== generator entry points vs the host emulator (opt-in cross-check)"
SSR_KERNEL_CROSSCHECK=1 timeout 600 python -m pytest tests/test_generator_entry_points_gpu.py -m gpu -q 2>&1 | tail -15
echo "== segmentation-regularised loss (opt-in, first GPU contact)"
SSR_ENABLE_SEG_LOSS=1 timeout 900 python -m pytest tests/test_seg_loss_gpu.py -m gpu -q 2>&1 | tail -40
echo "== generator only"
timeout 300 python scripts/gen_only.py --size 160 --iters 30
timeout 300 python scripts/gen_only.py --size 64 --iters 30 --defaults brain
echo "== generator launch list + ncu --set full (160^3, training() defaults)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02a_gen160_launches.csv \
    python scripts/gen_only.py --size 160 --iters 2 --warmup 2 > gpurun_out/r02a_gen160_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'deform_kernel|gmm_bias_kernel|blur3d|svf_step|resize_kernel|copy_strided|minmax' -s 30 -c 15 -f \
    -o gpurun_out/r02a_gen160 python scripts/gen_only.py --size 160 --iters 1 --warmup 2 > gpurun_out/r02a_gen160_ncu.log 2>&1
ls -la gpurun_out/r02a_gen160.ncu-rep
echo "== c4 (Hyperfine 192x192x64, 2 input channels)"
timeout 300 python scripts/c4_probe.py 2>&1 | tail -2
echo "== c5 (256^3)"
timeout 600 python bench.py --size 256 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench256.json 2> gpurun_out/r02a_bench256.err
cut -c1-400 gpurun_out/r02a_bench256.json; tail -3 gpurun_out/r02a_bench256.err
echo "== headline bench at HEAD"
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
cut -c1-400 gpurun_out/r02a_bench.json; tail -3 gpurun_out/r02a_bench.err
