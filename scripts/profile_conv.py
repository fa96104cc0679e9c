"""Runs single tcgen05 convolution launches at the headline sizes (for ncu captures and quick timing).
    python scripts/profile_conv.py [fwd24|fwd72|wgrad24|wgrad72|all] [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synthsr_b200._lib import lib, stream_ptr  # noqa: E402

CASES = {'fwd24': ('fwd', [160, 160, 160], 24, 0, 24), 'fwd72': ('fwd', [160, 160, 160], 24, 48, 24),
         'fwd48': ('fwd', [80, 80, 80], 48, 0, 48), 'fwd96': ('fwd', [40, 40, 40], 96, 0, 96),
         'fwd384': ('fwd', [10, 10, 10], 384, 0, 384), 'dgrad72': ('fwd', [160, 160, 160], 24, 0, 72),
         'wgrad24': ('wgrad', [160, 160, 160], 24, 0, 24), 'wgrad72': ('wgrad', [160, 160, 160], 24, 48, 24),
         'wgrad96': ('wgrad', [40, 40, 40], 96, 0, 96),
         # compensated forward of the same layers: 'x3' = bf16x3 (level 5), 'hy' = hybrid TF32 + bf16 (level 4)
         'fwd48x3': ('comp5', [80, 80, 80], 48, 0, 48), 'fwd48hy': ('comp4', [80, 80, 80], 48, 0, 48),
         'fwd96x3': ('comp5', [40, 40, 40], 96, 0, 96), 'fwd96hy': ('comp4', [40, 40, 40], 96, 0, 96),
         'fwd192x3': ('comp5', [20, 20, 20], 192, 0, 192), 'fwd192hy': ('comp4', [20, 20, 20], 192, 0, 192)}


def run(name, reps):
    kind, d, c1, c2, co = CASES[name]
    nv = int(np.prod(d))
    g = torch.Generator(device='cuda').manual_seed(0)
    x1 = torch.randn((nv, c1), device='cuda', generator=g)
    x2 = torch.randn((nv, c2), device='cuda', generator=g) if c2 else None
    if os.environ.get('SSR_CONST_DATA'):          # data-dependence experiment: constant operands
        x1 = torch.full_like(x1, float(os.environ['SSR_CONST_DATA']))
        x2 = torch.full_like(x2, float(os.environ['SSR_CONST_DATA'])) if c2 else None
    st = stream_ptr()
    flops = 2. * 27 * (c1 + c2) * co * nv
    if kind.startswith('comp'):
        level = int(kind[-1])
        w = torch.randn((3, 3, 3, c1, co), device='cuda', generator=g) / np.sqrt(27 * c1)
        b = torch.zeros(co, device='cuda')
        y = torch.empty((nv, co), device='cuda')
        xs = torch.empty((nv, 2 * c1), dtype=torch.bfloat16, device='cuda')
        pm = 9 if level == 5 else 7
        (lib.ssr_bf16x3_split if level == 5 else lib.ssr_tf32_split_bf16)(x1, xs, nv, c1, st)
        wp = torch.empty(lib.ssr_conv3d_packed_size(c1, c1, co, pm), device='cuda')
        lib.ssr_conv3d_pack_weights(w, wp, c1, c1, co, pm, st)
        fn = lambda: lib.ssr_conv3d_fwd_tc_comp(x1, xs, c1, wp, b, y, None, 1, *d, co, 1, 0, level, st)
    elif kind == 'fwd':
        w = torch.randn((3, 3, 3, c1 + c2, co), device='cuda', generator=g) / np.sqrt(27 * (c1 + c2))
        b = torch.zeros(co, device='cuda')
        y = torch.empty((nv, co), device='cuda')
        if c2 == 0 and c1 <= 32 and co <= 32 and not os.environ.get('SSR_NO_FWD_K2N'):
            wp = torch.empty(lib.ssr_conv3d_packed_size(c1, 0, co, 2), device='cuda')
            lib.ssr_conv3d_pack_weights(w, wp, c1, 0, co, 2, st)
            fn = lambda: lib.ssr_conv3d_fwd_tc_k2n(x1, c1, wp, b, y, 1, *d, co, 1, st)
        else:
            wp = torch.empty(lib.ssr_conv3d_packed_size(c1, c2, co, 0), device='cuda')
            lib.ssr_conv3d_pack_weights(w, wp, c1, c2, co, 0, st)
            fn = lambda: lib.ssr_conv3d_fwd_tc(x1, c1, x2, c2, wp, b, y, 1, *d, co, 1, st)
    else:
        dy = torch.randn((nv, co), device='cuda', generator=g)
        if os.environ.get('SSR_CONST_DATA'):
            dy = torch.full_like(dy, float(os.environ['SSR_CONST_DATA']))
        dw = torch.zeros(27 * (c1 + c2) * co, device='cuda')
        fn = lambda: lib.ssr_conv3d_wgrad_tc(x1, c1, x2, c2, dy, dw, None, None, 0, 1, *d, co, st)
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print('%-8s %s c=%d+%d->%d  %.3f ms  %.1f TFLOP/s' % (name, d, c1, c2, co, ms, flops / ms / 1e9))


def run_special(name, reps):
    """'up72': parity forward / gradient / weight gradient of the level-0 decoder convolution (48 upsampled channels of
    80^3 -> 24 channels at 160^3); 'first': the exact-fp32 first layer (1 -> 24 at 160^3) forward + weight gradient."""
    st = stream_ptr()
    g = torch.Generator(device='cuda').manual_seed(0)
    fns = {}
    if name == 'up72':
        dl, cs, cu, co = [80, 80, 80], 24, 48, 24
        nl, nf = 80 ** 3, 160 ** 3
        w = torch.randn((3, 3, 3, cs + cu, co), device='cuda', generator=g) / np.sqrt(27 * (cs + cu))
        low = torch.randn((nl, cu), device='cuda', generator=g)
        dy = torch.randn((nf, co), device='cuda', generator=g)
        wskip, weff = torch.empty(27 * cs * co, device='cuda'), torch.empty(8 * 27 * cu * co, device='cuda')
        lib.ssr_conv3d_up_weights(w, cs, cu, co, wskip, weff, st)
        nfw, ndg = lib.ssr_conv3d_packed_size(cu, 0, co, 0), lib.ssr_conv3d_packed_size(cu, 0, co, 1)
        fwd8, dgr8 = torch.empty(8 * nfw, device='cuda'), torch.empty(8 * ndg, device='cuda')
        for par in range(8):
            src = weff[par * 27 * cu * co:(par + 1) * 27 * cu * co]
            lib.ssr_conv3d_pack_weights(src, fwd8[par * nfw:(par + 1) * nfw], cu, 0, co, 0, st)
            lib.ssr_conv3d_pack_weights(src, dgr8[par * ndg:(par + 1) * ndg], cu, 0, co, 1, st)
        y, dlow = torch.empty((nf, co), device='cuda'), torch.empty((nl, cu), device='cuda')
        dw, scratch = torch.zeros((27, cs + cu, co), device='cuda'), torch.empty(8 * 27 * cu * co, device='cuda')
        fl = 2. * 27 * cu * co * nf
        fns = {'fwd_up': (lambda: lib.ssr_conv3d_fwd_tc_up(low, cu, fwd8, y, 1, *dl, co, st), fl),
               'dgrad_up': (lambda: lib.ssr_conv3d_dgrad_tc_up(dy, co, dgr8, dlow, 1, *dl, cu, st), fl),
               'wgrad_up': (lambda: lib.ssr_conv3d_wgrad_tc_up(low, cu, dy, dw, cs + cu, cs, scratch, 1, *dl, co, st), fl)}
    elif name == 'first':
        d, co = [160, 160, 160], 24
        nv = 160 ** 3
        x = torch.rand((nv, 1), device='cuda', generator=g)
        w = torch.randn((3, 3, 3, 1, co), device='cuda', generator=g) / np.sqrt(27)
        b = torch.zeros(co, device='cuda')
        y = torch.empty((nv, co), device='cuda')
        dy = torch.randn((nv, co), device='cuda', generator=g)
        dw = torch.zeros(27 * co, device='cuda')
        fl = 2. * 27 * co * nv
        fns = {'first_fwd': (lambda: lib.ssr_conv3d_fwd_ref(x, 1, None, 0, w, b, y, 1, *d, co, 3, 1, st), fl),
               'first_wgrad': (lambda: lib.ssr_conv3d_wgrad_ref(x, 1, None, 0, dy, dw, None, 1, *d, co, 3, st), fl)}
    for k, (fn, fl) in fns.items():
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print('%-12s %.3f ms  %.1f TFLOP/s (algorithmic)' % (k, ms, fl / ms / 1e9))


if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'all'
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    for n in (CASES if which == 'all' else which.split(',')):
        if n in ('up72', 'first'):
            run_special(n, reps)
        else:
            run(n, reps)
