import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synthsr_b200._lib import lib, stream_ptr
def run(N, nacc, chain, commit, cyc, nblk=148):
    out = torch.zeros(nblk, device='cuda')
    lib.ssr_tc_microbench(out, nblk, N, nacc, chain, 4096, 1, commit, cyc, stream_ptr())
    torch.cuda.synchronize()
    return out.mean().item()
for N in (32,):
    for cyc in (0, 1):
        print('N=%d cycle_addr=%d: no-commit %.1f | commit/27 %.1f | commit/9 %.1f | commit/3 %.1f | 3acc chain9 commit/27 %.1f' % (
            N, cyc, run(N, 1, 1, 0, cyc), run(N, 1, 1, 27, cyc), run(N, 1, 1, 9, cyc), run(N, 1, 1, 3, cyc), run(N, 3, 9, 27, cyc)))
print('N=80: no-commit %.1f commit/9 %.1f | N=96 no-commit %.1f commit/12 %.1f' % (run(80, 1, 1, 0, 0), run(80, 1, 1, 9, 0), run(96, 1, 1, 0, 0), run(96, 1, 1, 12, 0)))
