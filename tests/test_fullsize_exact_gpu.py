"""Full-size (160^3, BASELINE config) checks of the tensor-core convolutions through exactly representable data:
small-integer activations and weights make every product exact in TF32 and every partial sum an integer below 2^24, so
the expected result is known in closed form (box sums / tap counts) and the comparison is BIT-EXACT.  Covers every tile,
d0 range, ring slot and channel part of the kernels at the size the benchmark runs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _box_sum(a):
    """3x3x3 'same' zero-padded box sum of an integer volume."""
    out = np.zeros_like(a)
    D = a.shape
    for k0 in (-1, 0, 1):
        for k1 in (-1, 0, 1):
            for k2 in (-1, 0, 1):
                src = a[max(k0, 0):D[0] + min(k0, 0), max(k1, 0):D[1] + min(k1, 0), max(k2, 0):D[2] + min(k2, 0)]
                out[max(-k0, 0):D[0] + min(-k0, 0), max(-k1, 0):D[1] + min(-k1, 0), max(-k2, 0):D[2] + min(-k2, 0)] += src
    return out


def _pattern(d):
    i0, i1, i2 = np.meshgrid(np.arange(d[0]), np.arange(d[1]), np.arange(d[2]), indexing='ij')
    return (i0 % 5 + 2 * (i2 % 3) + (i1 % 2)).astype(np.int64)        # 0..9


@pytest.mark.parametrize('d,c1,c2,co', [([160, 160, 160], 24, 0, 24), ([160, 160, 160], 24, 48, 24),
                                        ([80, 80, 80], 48, 0, 48), ([10, 10, 10], 384, 0, 384)])
def test_forward_exact_at_full_size(d, c1, c2, co):
    from synthsr_b200._lib import lib, stream_ptr
    nv = int(np.prod(d))
    base = _pattern(d)
    cin = c1 + c2
    cpar = (np.arange(cin) % 2).astype(np.int64)                       # x[v][c] = base[v] + c % 2
    x = torch.from_numpy((base.reshape(-1, 1) + cpar[None, :]).astype(np.float32)).cuda()
    x1 = x[:, :c1].contiguous()
    x2 = x[:, c1:].contiguous() if c2 else None
    wscale = 1 + (np.arange(co) % 2)                                    # w[.., ci, co] = 1 + co % 2
    w = torch.from_numpy(np.broadcast_to(wscale.astype(np.float32), (3, 3, 3, cin, co)).copy()).cuda()
    b = torch.zeros(co, device='cuda')
    expect = _box_sum(cin * base + int(cpar.sum()))                     # sum over channels, then over the 27 taps
    expect = torch.from_numpy(expect.reshape(-1, 1) * wscale[None, :]).float()
    assert expect.max().item() < 2 ** 24
    st = stream_ptr()
    y = torch.full((nv, co), float('nan'), device='cuda')
    wp = torch.empty(lib.ssr_conv3d_packed_size(c1, c2, co, 0), device='cuda')
    lib.ssr_conv3d_pack_weights(w, wp, c1, c2, co, 0, st)
    lib.ssr_conv3d_fwd_tc(x1, c1, x2, c2, wp, b, y, 1, *d, co, 0, st)
    torch.cuda.synchronize()
    assert torch.equal(y.cpu(), expect), 'generic kernel'
    if co <= 32 and c1 <= 32:                                            # d2-taps-in-N kernel, whole or in channel parts
        y.fill_(float('nan'))
        parts = [(x1, c1, 0, c1, 0)] + ([(x2, c2, o, min(32, c2 - o), c1 + o) for o in range(0, c2, 32)] if c2 else [])
        keep = []
        for i, (src, ctot, c0, cn, coff) in enumerate(parts):
            wq = torch.empty(lib.ssr_conv3d_packed_size(cin, (coff << 8) | cn, co, 4), device='cuda')
            lib.ssr_conv3d_pack_weights(w, wq, cin, (coff << 8) | cn, co, 4, st)
            keep.append(wq)
            lib.ssr_conv3d_fwd_tc_k2n_part(src, ctot, c0, cn, wq, b, y, 1, *d, co, 0, 1 if i else 0,
                                           1 if i == len(parts) - 1 else 0, st)
        torch.cuda.synchronize()
        assert torch.equal(y.cpu(), expect), 'k2n kernel'


@pytest.mark.parametrize('d,c1,c2,co', [([160, 160, 160], 24, 0, 24), ([160, 160, 160], 24, 48, 24),
                                        ([40, 40, 40], 96, 0, 96), ([10, 10, 10], 192, 0, 384)])
def test_weight_gradient_exact_at_full_size(d, c1, c2, co):
    """x = ci % 2 + 1 on every voxel, dy = co % 3 on every voxel: dW[k][ci][co] = x * dy * #(voxels whose tap k is inside)."""
    from synthsr_b200._lib import lib, stream_ptr
    nv = int(np.prod(d))
    cin = c1 + c2
    xv = (np.arange(cin) % 2 + 1).astype(np.float32)
    dv = (np.arange(co) % 3).astype(np.float32)
    x = torch.from_numpy(np.broadcast_to(xv, (nv, cin)).copy()).cuda()
    dy = torch.from_numpy(np.broadcast_to(dv, (nv, co)).copy()).cuda()
    x1 = x[:, :c1].contiguous()
    x2 = x[:, c1:].contiguous() if c2 else None
    cnt = np.zeros((3, 3, 3), dtype=np.int64)
    for k0 in range(3):
        for k1 in range(3):
            for k2 in range(3):
                cnt[k0, k1, k2] = (d[0] - (k0 != 1)) * (d[1] - (k1 != 1)) * (d[2] - (k2 != 1))
    expect = cnt[:, :, :, None, None] * xv[None, None, None, :, None].astype(np.int64) * dv[None, None, None, None, :].astype(np.int64)
    assert expect.max() < 2 ** 24
    dw = torch.zeros(27 * cin * co, device='cuda')
    lib.ssr_conv3d_wgrad_tc(x1, c1, x2, c2, dy, dw, None, None, 0, 1, *d, co, stream_ptr())
    torch.cuda.synchronize()
    assert torch.equal(dw.cpu(), torch.from_numpy(expect.astype(np.float32).reshape(-1)))
