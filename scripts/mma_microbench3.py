"""tf32 MMA cost with the kernels' own asm-chained issue: MN-major (SWIZZLE_128B_BASE32B) chains of 16 vs K-major chains of 4.
148 CTAs.  GPU box only."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synthsr_b200._lib import lib, stream_ptr
for mode, name in ((3, 'K-major  chain4 '), (2, 'MN-major chain16')):
    for N in (32, 48, 64, 96, 128, 160, 192, 256):
        for nacc in (1, 3):
            if nacc * N > 512:
                continue
            out = torch.zeros(148, device='cuda')
            lib.ssr_tc_microbench(out, 148, N, nacc, 16, 4096, mode, 0, 0, stream_ptr())
            torch.cuda.synchronize()
            print('%s N=%3d nacc=%d  %.1f cycles/MMA' % (name, N, nacc, out.mean().item()))
