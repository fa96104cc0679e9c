"""do warps parked on an mbarrier (try_wait + suspend hint) slow the tensor pipe?  148 CTAs.  GPU box only."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synthsr_b200._lib import lib, stream_ptr
for mode, name, N in ((2, 'MN-major chain16 N=96', 96), (3, 'K-major chain4 N=32', 32), (3, 'K-major chain4 N=96', 96)):
    for ce in (0, 3):
        for npoll in (0, 1, 2, 3):
            out = torch.zeros(148, device='cuda')
            lib.ssr_tc_microbench(out, 148, N, 3, 16, 4096 * 2, mode, ce, npoll, stream_ptr())
            torch.cuda.synchronize()
            print('%s  commit every %d chains, %d parked warps  %.1f cycles/MMA' % (name, ce, npoll, out.mean().item()))
