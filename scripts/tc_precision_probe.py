"""Measures the TF32 tensor-core convolution error against a float64 reference for the operand-rounding variants
(run on the GPU box):  python scripts/tc_precision_probe.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synthsr_b200._lib import lib, stream_ptr  # noqa: E402


def rna_tf32(t):
    u = t.view(torch.int32)
    u = (u + 0x1000) & ~0x1FFF          # round-to-nearest (ties away) on the 13 dropped bits
    return u.view(torch.float32)


def run(d, c1, co, round_act, tma_tf32, pack_trunc):
    rng = np.random.default_rng(0)
    nv = int(np.prod(d))
    x = torch.from_numpy(rng.normal(size=(nv, c1)).astype(np.float32)).cuda()
    w = torch.from_numpy((rng.normal(size=(3, 3, 3, c1, co)) / np.sqrt(27 * c1)).astype(np.float32)).cuda()
    ref = torch.nn.functional.conv3d(x.double().view(1, *d, c1).permute(0, 4, 1, 2, 3),
                                     w.double().permute(4, 3, 0, 1, 2), padding=1).permute(0, 2, 3, 4, 1).reshape(nv, co)
    os.environ.pop('SSR_TMA_DTYPE', None)
    os.environ.pop('SSR_PACK_TRUNC', None)
    if tma_tf32:
        os.environ['SSR_TMA_DTYPE'] = 'tf32'
    if pack_trunc:
        os.environ['SSR_PACK_TRUNC'] = '1'
    xin = rna_tf32(x.clone()) if round_act else x
    y = torch.empty((nv, co), dtype=torch.float32, device='cuda')
    wp = torch.empty(lib.ssr_conv3d_packed_size(c1, 0, co, 0), dtype=torch.float32, device='cuda')
    st = stream_ptr()
    lib.ssr_conv3d_pack_weights(w, wp, c1, 0, co, 0, st)
    lib.ssr_conv3d_fwd_tc(xin, c1, None, 0, wp, None, y, 1, *d, co, 0, st)
    torch.cuda.synchronize()
    e = (y.double() - ref)
    return e.abs().max().item() / ref.abs().max().item(), e.norm().item() / ref.norm().item(), (e.mean() / ref.abs().mean()).item()


if __name__ == '__main__':
    for (d, c1, co) in [([16, 16, 16], 24, 24), ([16, 16, 16], 96, 96)]:
        for name, kw in [('trunc/trunc (raw fp32 operands)', dict(round_act=0, tma_tf32=0, pack_trunc=1)),
                         ('act trunc, weights RN', dict(round_act=0, tma_tf32=0, pack_trunc=0)),
                         ('act RN (pre-rounded), weights RN', dict(round_act=1, tma_tf32=0, pack_trunc=0)),
                         ('TMA TFLOAT32 maps, weights RN', dict(round_act=0, tma_tf32=1, pack_trunc=0))]:
            mx, l2, bias = run(d, c1, co, **kw)
            print('%-12s %-36s max/max %.2e  relL2 %.2e  mean-bias %.2e' % ('%d->%d' % (c1, co), name, mx, l2, bias))
