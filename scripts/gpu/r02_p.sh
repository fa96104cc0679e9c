#!/bin/bash
# round 2, call P: validation of the final state (scripts/gpu/validate.sh) + kernel-level evidence for the bf16x3 scheme
bash scripts/gpu/validate.sh r02p
echo "== single launches: plain TF32 / hybrid / bf16x3"
timeout 300 python scripts/profile_conv.py fwd48,fwd48hy,fwd48x3,fwd96,fwd96hy,fwd96x3,fwd192hy,fwd192x3 20 2>&1 | tail -8 | tee gpurun_out/r02p_conv_schemes.txt
echo "== ncu --set full: 48 -> 48 @ 80^3, the three schemes"
for c in fwd48 fwd48hy fwd48x3; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3d_tc_kernel -s 1 -c 1 -f -o gpurun_out/r02p_$c \
      python scripts/profile_conv.py $c 2 > gpurun_out/r02p_ncu_$c.log 2>&1
  ls -la gpurun_out/r02p_$c.ncu-rep 2>&1 | cut -c20-
done
