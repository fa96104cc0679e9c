"""does the operand placement in shared memory / run length change the MN-major N=96 MMA rate?  GPU box only."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synthsr_b200._lib import lib, stream_ptr
for iters in (8192, 65536):
    for b_kb, nslab in ((72, 3), (90, 5), (90, 3), (72, 5), (96, 5), (100, 4)):
        out = torch.zeros(148, device='cuda')
        lib.ssr_tc_microbench(out, 148, 96, 3, b_kb, iters, 2, nslab, 0, stream_ptr())
        torch.cuda.synchronize()
        print('iters %6d  B at %3d KB, %d A slabs  %.1f cycles/MMA (min %.1f max %.1f)' % (iters, b_kb, nslab, out.mean().item(), out.min().item(), out.max().item()))
