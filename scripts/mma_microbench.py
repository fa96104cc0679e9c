"""cycles per tcgen05.mma (M=128, K=8 tf32, SS) vs N / accumulator switching.  Run on the GPU box."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synthsr_b200._lib import lib, stream_ptr
for nblk in (1, 148):
    for kmajor in (1, 0):
        for N in (16, 32, 48, 64, 96, 128, 192, 256):
            row = []
            for nacc, chain in ((1, 1), (2, 1), (4, 1), (2, 9), (2, 27)):
                if nacc * N > 512:
                    row.append('   -  '); continue
                out = torch.zeros(nblk, device='cuda')
                lib.ssr_tc_microbench(out, nblk, N, nacc, chain, 4096, kmajor, stream_ptr())
                torch.cuda.synchronize()
                row.append('%6.1f' % out.mean().item())
            print('blocks=%3d %s N=%3d  cyc/MMA: same-acc %s | alt2 %s | alt4 %s | 2acc chain9 %s | 2acc chain27 %s  (tensor floor %d)' % (
                nblk, 'K-major ' if kmajor else 'MN-major', N, *row, N // 2))
