"""tcgen05.mma issue-cost microbenchmarks (cycles per MMA per CTA, ssr_tc_microbench); results in profiles/r01_mma_*.txt.
GPU box only.     python scripts/mma_microbench.py [shape|commit|chains|parked|placement|all]

  shape      cycles per MMA (M=128, K=8 tf32, operands from shared memory) vs N and accumulator switching
  commit     cost of tcgen05.commit every n MMAs, cycle-counter addressing
  chains     the kernels' own asm-chained issue: MN-major chains of 16 vs K-major chains of 4
  parked     do warps parked on an mbarrier (try_wait + suspend hint) slow the tensor pipe?
  placement  does the operand placement in shared memory / run length change the MN-major N=96 rate?
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synthsr_b200._lib import lib, stream_ptr  # noqa: E402


def run(N, nacc, chain, iters, mode, commit=0, cyc=0, nblk=148):
    out = torch.zeros(nblk, device='cuda')
    lib.ssr_tc_microbench(out, nblk, N, nacc, chain, iters, mode, commit, cyc, stream_ptr())
    torch.cuda.synchronize()
    return out


def shape():
    for nblk in (1, 148):
        for kmajor in (1, 0):
            for N in (16, 32, 48, 64, 96, 128, 192, 256):
                row = []
                for nacc, chain in ((1, 1), (2, 1), (4, 1), (2, 9), (2, 27)):
                    row.append('   -  ' if nacc * N > 512 else '%6.1f' % run(N, nacc, chain, 4096, kmajor, nblk=nblk).mean().item())
                print('blocks=%3d %s N=%3d  cyc/MMA: same-acc %s | alt2 %s | alt4 %s | 2acc chain9 %s | 2acc chain27 %s  (tensor floor %d)' % (
                    nblk, 'K-major ' if kmajor else 'MN-major', N, *row, N // 2))


def commit():
    r = lambda N, nacc, chain, c, cyc: run(N, nacc, chain, 4096, 1, c, cyc).mean().item()
    for cyc in (0, 1):
        print('N=32 cycle_addr=%d: no-commit %.1f | commit/27 %.1f | commit/9 %.1f | commit/3 %.1f | 3acc chain9 commit/27 %.1f' % (
            cyc, r(32, 1, 1, 0, cyc), r(32, 1, 1, 27, cyc), r(32, 1, 1, 9, cyc), r(32, 1, 1, 3, cyc), r(32, 3, 9, 27, cyc)))
    print('N=80: no-commit %.1f commit/9 %.1f | N=96 no-commit %.1f commit/12 %.1f' % (
        r(80, 1, 1, 0, 0), r(80, 1, 1, 9, 0), r(96, 1, 1, 0, 0), r(96, 1, 1, 12, 0)))


def chains():
    for mode, name in ((3, 'K-major  chain4 '), (2, 'MN-major chain16')):
        for N in (32, 48, 64, 96, 128, 160, 192, 256):
            for nacc in (1, 3):
                if nacc * N <= 512:
                    print('%s N=%3d nacc=%d  %.1f cycles/MMA' % (name, N, nacc, run(N, nacc, 16, 4096, mode).mean().item()))


def parked():
    for mode, name, N in ((2, 'MN-major chain16 N=96', 96), (3, 'K-major chain4 N=32', 32), (3, 'K-major chain4 N=96', 96)):
        for ce in (0, 3):
            for npoll in (0, 1, 2, 3):
                print('%s  commit every %d chains, %d parked warps  %.1f cycles/MMA' % (
                    name, ce, npoll, run(N, 3, 16, 8192, mode, ce, npoll).mean().item()))


def placement():
    for iters in (8192, 65536):
        for b_kb, nslab in ((72, 3), (90, 5), (90, 3), (72, 5), (96, 5), (100, 4)):
            out = run(96, 3, b_kb, iters, 2, nslab, 0)
            print('iters %6d  B at %3d KB, %d A slabs  %.1f cycles/MMA (min %.1f max %.1f)' % (
                iters, b_kb, nslab, out.mean().item(), out.min().item(), out.max().item()))


if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'all'
    for name, fn in (('shape', shape), ('commit', commit), ('chains', chains), ('parked', parked), ('placement', placement)):
        if which in (name, 'all'):
            print('== ' + name)
            fn()
