"""Parity of the BENCHMARKED convolution mode (conv_impl='tc3': tcgen05, forward compensated to fp32-class accuracy,
backward plain TF32) against the float64 oracle at north_star's bars, un-widened:

    prediction: relative L2 <= 1e-3 AND max|err| / max|pred| <= 1e-3;  loss <= 1e-3;  every gradient tensor <= 1e-2

* kernel level: every compensated entry point against a float64 convolution of the UNROUNDED operands (what KL.Conv3D
  computes in fp32, ext/neuron/models.py:316,444,481);
* full training step, reference topology (24 features, 5 levels), random init, against oracle/unet.py in float64 on the
  CPU: uniform noise at 32^3 / 64^3 (l1, l2: every bar, every tensor) and at 96^3 (noise l1 / l2, and a batch of the
  benchmark's own distribution: label phantom -> CUDA generator); the reference's trained weights on a crop of the
  reference's scan (every bar, every tensor);
* 160^3 (BASELINE configs[1], the benchmark size), generated batch and noise: against the exact-fp32 CUDA-core mode
  (conv_impl='ref', itself within 2e-5 of the float64 oracle where the CPU oracle reaches), because a float64 CPU step at
  160^3 needs ~30 GB.
* gradients: EVERY tensor is gated at 1e-2 wherever the float64 oracle runs (<= 96^3), with the oracle's MaxPooling3D taking
  the same window winners as the GPU forward (oracle.unet._maxpool_routed; the imposed winner must be within 5e-5 of the
  float64 window maximum in every window).  Under free routing a single near-tied window resolved differently by
  two non-bit-identical forwards moves a whole level's gradient by sqrt(2 / #windows) -- the EXACT-fp32 mode is 5e-3 off
  float64 at 96^3 that way (test_gradient_comparison_is_limited_by_maxpool_argmax_flips, profiles/r02_actgrad_96_noise_l2.txt).
  At 160^3 (no float64 reference) the whole gradient vector is gated and the individual tensors are recorded.

Every measured number is appended to gpurun_out/unet_parity.txt.  scripts/tf32_error_emulation.py reproduces the error
levels on the CPU and is how the set of compensated layers was chosen."""
import os

import numpy as np
import pytest
import torch

from helpers import gpu_pool_routing as _gpu_routing

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRED_TOL, LOSS_TOL, GRAD_TOL = 1e-3, 1e-3, 1e-2          # north_star
# one compensated convolution against float64: what remains is the fp32 accumulation of up to 3 x 27 x 384 products in TMEM
# (measured 1.3e-5 at K = 3 x 1296, 4.2e-5 at K = 3 x 5184); plain TF32 sits at 3e-4 .. 2e-3 on the same inputs
KERNEL_TOL = 1e-4
WELL_POSED = 1e-3      # what the exact-fp32 mode reaches on every gradient tensor once the pooling winners agree


def _log(line):
    try:
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        with open(os.path.join(ROOT, 'gpurun_out', 'unet_parity.txt'), 'a') as f:
            f.write(line + '\n')
    except OSError:
        pass


def _find(rel):
    for base in (os.path.join(ROOT, 'baseline', '_ref'), '/root/reference', '/root/reference/data'):
        p = os.path.join(base, rel)
        if os.path.isfile(p):
            return p
    return None


def _conv64(x, w, b, d, elu=True):
    """float64 3x3x3 'same' convolution of [nv, C] activations on grid d with a (3,3,3,Cin,Cout) kernel."""
    xr = x.double().cpu().view(1, *d, -1).permute(0, 4, 1, 2, 3)
    wr = w.double().cpu().permute(4, 3, 0, 1, 2)
    y = torch.nn.functional.conv3d(xr, wr, None if b is None else b.double().cpu(), padding=1)
    if elu:
        y = torch.nn.functional.elu(y)
    return y.permute(0, 2, 3, 4, 1).reshape(-1, w.shape[-1])


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def test_tf32_residual_is_the_part_the_tma_rounds_away():
    from synthsr_b200._lib import lib, stream_ptr
    rng = np.random.default_rng(0)
    x = _t(rng.normal(size=100003) * np.exp(rng.normal(size=100003) * 3))
    lo = torch.full_like(x, float('nan'))
    lib.ssr_tf32_residual(x, lo, x.numel(), stream_ptr())
    torch.cuda.synchronize()
    u = x.view(torch.int32)
    hi = ((u + 0xFFF + ((u >> 13) & 1)) & ~0x1FFF).view(torch.float32)      # round to nearest even TF32
    assert torch.equal(lo, x - hi)
    assert torch.equal(hi + lo, x)                                            # the split is exact
    assert (lo.abs() <= x.abs() * 2.0 ** -11).all()


@pytest.mark.parametrize('d,c,co,level', [([8, 16, 24], 48, 96, 3), ([10, 10, 10], 192, 384, 3), ([16, 16, 16], 24, 48, 3),
                                          ([20, 20, 20], 96, 192, 3), ([8, 16, 24], 48, 96, 2), ([12, 20, 8], 384, 384, 3)])
def test_compensated_generic_forward_matches_float64(d, c, co, level):
    """ssr_conv3d_fwd_tc_comp (with and without the BatchNorm sums): level 3 reaches fp32 accuracy; level 2 leaves the
    weight rounding (compared with a float64 convolution of the rna-rounded weights)."""
    from synthsr_b200._lib import lib, stream_ptr
    rng = np.random.default_rng(1)
    nv = int(np.prod(d))
    x = _t(rng.normal(size=(nv, c)))
    w = _t(rng.normal(size=(3, 3, 3, c, co)) / np.sqrt(27 * c))
    b = _t(rng.normal(size=co))
    st = stream_ptr()
    lo = torch.empty_like(x)
    lib.ssr_tf32_residual(x, lo, x.numel(), st)
    wp = torch.empty(lib.ssr_conv3d_packed_size(c, c, co, 5), dtype=torch.float32, device='cuda')
    lib.ssr_conv3d_pack_weights(w, wp, c, c, co, 5, st)
    wref = w
    if level == 2:
        uw = w.view(torch.int32)
        wref = ((uw + 0x1000) & ~0x1FFF).view(torch.float32)
    y64 = _conv64(x, wref, b, d)
    for with_sums in (False, True):
        y = torch.full((nv, co), float('nan'), dtype=torch.float32, device='cuda')
        sums = torch.full((2 * co,), float('nan'), dtype=torch.float64, device='cuda') if with_sums else None
        lib.ssr_conv3d_fwd_tc_comp(x, lo, c, wp, b, y, sums, 1, *d, co, 1, 0, level, st)
        torch.cuda.synchronize()
        err = (y.double().cpu() - y64).abs().max().item() / y64.abs().max().item()
        _log('comp generic level %d %s %d->%d sums=%d: max/max %.2e' % (level, d, c, co, with_sums, err))
        # 3xTF32 (the cross-check scheme) accumulates three K = 27 x 384 chains in fp32: 9.2e-5 on the longest case; bound with
        # margin there, KERNEL_TOL elsewhere
        assert err < (1.5 * KERNEL_TOL if c >= 384 else KERNEL_TOL), (d, c, co, level, err)
        if with_sums:
            s = sums.cpu().numpy()
            assert np.allclose(s[:co], y.double().sum(0).cpu().numpy(), rtol=1e-6, atol=1e-6 * nv)
            assert np.allclose(s[co:], (y.double() ** 2).sum(0).cpu().numpy(), rtol=1e-6, atol=1e-6 * nv)


def test_compensated_generic_forward_accumulates_channel_parts():
    """a concatenated input [x1, x2] as two compensated launches (second one accumulates, + bias + ELU)."""
    from synthsr_b200._lib import lib, stream_ptr
    rng = np.random.default_rng(2)
    d, c1, c2, co = [8, 16, 16], 48, 96, 48
    nv = int(np.prod(d))
    x1, x2 = _t(rng.normal(size=(nv, c1))), _t(rng.normal(size=(nv, c2)))
    w = _t(rng.normal(size=(3, 3, 3, c1 + c2, co)) / np.sqrt(27 * (c1 + c2)))
    b = _t(rng.normal(size=co))
    st = stream_ptr()
    y = torch.full((nv, co), float('nan'), dtype=torch.float32, device='cuda')
    lo = torch.empty(nv * max(c1, c2), dtype=torch.float32, device='cuda')
    for i, (x, c, enc) in enumerate(((x1, c1, c1), (x2, c2, (c1 << 12) | c2))):
        wp = torch.empty(lib.ssr_conv3d_packed_size(c1 + c2, enc, co, 5), dtype=torch.float32, device='cuda')
        lib.ssr_conv3d_pack_weights(w, wp, c1 + c2, enc, co, 5, st)
        lib.ssr_tf32_residual(x, lo, x.numel(), st)
        lib.ssr_conv3d_fwd_tc_comp(x, lo, c, wp, b if i else None, y, None, 1, *d, co, i, i, 3, st)
    torch.cuda.synchronize()
    y64 = _conv64(torch.cat([x1, x2], 1), w, b, d)
    err = (y.double().cpu() - y64).abs().max().item() / y64.abs().max().item()
    assert err < KERNEL_TOL, err


@pytest.mark.parametrize('d', [[16, 16, 32], [12, 20, 18]])
def test_compensated_k2n_forward_matches_float64(d):
    """the three k2n passes of a 24 -> 24 full-resolution layer (hi.hi, lo.hi, hi.lo; the last one with the BatchNorm sums)."""
    from synthsr_b200._lib import lib, stream_ptr
    rng = np.random.default_rng(3)
    c = co = 24
    nv = int(np.prod(d))
    x = _t(rng.normal(size=(nv, c)))
    w = _t(rng.normal(size=(3, 3, 3, c, co)) / np.sqrt(27 * c))
    b = _t(rng.normal(size=co))
    st = stream_ptr()
    lo = torch.empty_like(x)
    lib.ssr_tf32_residual(x, lo, x.numel(), st)
    whi = torch.empty(lib.ssr_conv3d_packed_size(c, 0, co, 2), dtype=torch.float32, device='cuda')
    wlo = torch.empty(lib.ssr_conv3d_packed_size(c, 0, co, 6), dtype=torch.float32, device='cuda')
    lib.ssr_conv3d_pack_weights(w, whi, c, 0, co, 2, st)
    lib.ssr_conv3d_pack_weights(w, wlo, c, 0, co, 6, st)
    y = torch.full((nv, co), float('nan'), dtype=torch.float32, device='cuda')
    sums = torch.full((2 * co,), float('nan'), dtype=torch.float64, device='cuda')
    lib.ssr_conv3d_fwd_tc_k2n_part(x, c, 0, c, whi, b, y, 1, *d, co, 1, 0, 0, st)
    lib.ssr_conv3d_fwd_tc_k2n_part(lo, c, 0, c, whi, b, y, 1, *d, co, 1, 1, 0, st)
    lib.ssr_conv3d_fwd_tc_k2n_part_stats(x, c, 0, c, wlo, b, y, sums, 1, *d, co, 1, 1, st)
    torch.cuda.synchronize()
    y64 = _conv64(x, w, b, d)
    err = (y.double().cpu() - y64).abs().max().item() / y64.abs().max().item()
    _log('comp k2n %s: max/max %.2e' % (d, err))
    assert err < KERNEL_TOL, err
    s = sums.cpu().numpy()
    assert np.allclose(s[:co], y.double().sum(0).cpu().numpy(), rtol=1e-6, atol=1e-6 * nv)
    assert np.allclose(s[co:], (y.double() ** 2).sum(0).cpu().numpy(), rtol=1e-6, atol=1e-6 * nv)


@pytest.mark.parametrize('dl,cu,co', [([8, 8, 16], 96, 48), ([10, 12, 8], 192, 96)])
def test_compensated_parity_forward_matches_float64(dl, cu, co):
    """ssr_conv3d_fwd_tc_up_comp: convolution over the 2x nearest-upsampled tensor from the low-resolution tensor."""
    from synthsr_b200._lib import lib, stream_ptr
    rng = np.random.default_rng(4)
    cs = co
    nl = int(np.prod(dl))
    df = [2 * v for v in dl]
    low = _t(rng.normal(size=(nl, cu)))
    w = _t(rng.normal(size=(3, 3, 3, cs + cu, co)) / np.sqrt(27 * (cs + cu)))
    st = stream_ptr()
    wskip = torch.empty(27 * cs * co, dtype=torch.float32, device='cuda')
    weff = torch.empty(8 * 27 * cu * co, dtype=torch.float32, device='cuda')
    lib.ssr_conv3d_up_weights(w, cs, cu, co, wskip, weff, st)
    n5 = lib.ssr_conv3d_packed_size(cu, cu, co, 5)
    wp8 = torch.empty(8 * n5, dtype=torch.float32, device='cuda')
    for par in range(8):
        lib.ssr_conv3d_pack_weights(weff[par * 27 * cu * co:], wp8[par * n5:], cu, cu, co, 5, st)
    lo = torch.empty_like(low)
    lib.ssr_tf32_residual(low, lo, low.numel(), st)
    y = torch.full((8 * nl, co), float('nan'), dtype=torch.float32, device='cuda')
    lib.ssr_conv3d_fwd_tc_up_comp(low, lo, cu, wp8, y, 1, *dl, co, 3, st)
    torch.cuda.synchronize()
    up = low.view(*dl, cu).repeat_interleave(2, 0).repeat_interleave(2, 1).repeat_interleave(2, 2).reshape(-1, cu)
    y64 = _conv64(up, w[:, :, :, cs:, :], None, df, elu=False)
    err = (y.double().cpu() - y64).abs().max().item() / y64.abs().max().item()
    _log('comp parity %s %d->%d: max/max %.2e' % (dl, cu, co, err))
    assert err < KERNEL_TOL, err


def test_tf32_split_bf16_layout():
    """x2 = [bf16(x - rne_tf32(x)) | bf16(rne_tf32(x))], 2C bf16 channels per voxel"""
    from synthsr_b200._lib import lib, stream_ptr
    rng = np.random.default_rng(0)
    nv, c = 1000, 24
    x = _t(rng.normal(size=(nv, c)) * np.exp(rng.normal(size=(nv, c)) * 2))
    x2 = torch.zeros((nv, 2 * c), dtype=torch.bfloat16, device='cuda')
    lib.ssr_tf32_split_bf16(x, x2, nv, c, stream_ptr())
    torch.cuda.synchronize()
    u = x.view(torch.int32)
    hi = ((u + 0xFFF + ((u >> 13) & 1)) & ~0x1FFF).view(torch.float32)
    assert torch.equal(x2[:, c:], hi.to(torch.bfloat16)) and torch.equal(x2[:, :c], (x - hi).to(torch.bfloat16))


_SCHEME = {4: ('ssr_tf32_split_bf16', 7), 5: ('ssr_bf16x3_split', 9)}      # level -> (activation split, weight pack mode)


@pytest.mark.parametrize('d,c,co', [([8, 16, 24], 48, 96), ([10, 10, 10], 192, 384), ([16, 16, 16], 24, 48),
                                    ([20, 20, 20], 96, 192), ([12, 20, 8], 384, 384), ([16, 16, 16], 40, 24)])
@pytest.mark.parametrize('level', [4, 5])
def test_hybrid_generic_forward_matches_float64(d, c, co, level):
    """level 4 of ssr_conv3d_fwd_tc_comp: TF32 main term + ONE bf16 chain for both correction terms;
    level 5 (bf16x3): x1 w1 + x2 w1 + x1 w2 with 8-bit pieces, every term a bf16 MMA"""
    from synthsr_b200._lib import lib, stream_ptr
    split, pm = _SCHEME[level]
    rng = np.random.default_rng(1)
    nv = int(np.prod(d))
    x = _t(rng.normal(size=(nv, c)))
    w = _t(rng.normal(size=(3, 3, 3, c, co)) / np.sqrt(27 * c))
    b = _t(rng.normal(size=co))
    st = stream_ptr()
    x2 = torch.empty((nv, 2 * c), dtype=torch.bfloat16, device='cuda')
    getattr(lib, split)(x, x2, nv, c, st)
    wp = torch.empty(lib.ssr_conv3d_packed_size(c, c, co, pm), dtype=torch.float32, device='cuda')
    lib.ssr_conv3d_pack_weights(w, wp, c, c, co, pm, st)
    y64 = _conv64(x, w, b, d)
    for with_sums in (False, True):
        y = torch.full((nv, co), float('nan'), dtype=torch.float32, device='cuda')
        sums = torch.full((2 * co,), float('nan'), dtype=torch.float64, device='cuda') if with_sums else None
        lib.ssr_conv3d_fwd_tc_comp(x, x2, c, wp, b, y, sums, 1, *d, co, 1, 0, level, st)
        torch.cuda.synchronize()
        err = (y.double().cpu() - y64).abs().max().item() / y64.abs().max().item()
        _log('%s generic %s %d->%d sums=%d: max/max %.2e' % (split, d, c, co, with_sums, err))
        assert err < KERNEL_TOL, (d, c, co, err)
        if with_sums:
            s = sums.cpu().numpy()
            assert np.allclose(s[:co], y.double().sum(0).cpu().numpy(), rtol=1e-6, atol=1e-6 * nv)


@pytest.mark.parametrize('d', [[16, 16, 32], [12, 20, 18]])
def test_hybrid_k2n_forward_matches_float64(d):
    """24 -> 24 full-resolution layer: TF32 pass (x, w_hi) + bf16 pass ([x_lo | x_hi], [w_hi ; w_lo]) with the BatchNorm sums"""
    from synthsr_b200._lib import lib, stream_ptr
    rng = np.random.default_rng(3)
    c = co = 24
    nv = int(np.prod(d))
    x = _t(rng.normal(size=(nv, c)))
    w = _t(rng.normal(size=(3, 3, 3, c, co)) / np.sqrt(27 * c))
    b = _t(rng.normal(size=co))
    st = stream_ptr()
    x2 = torch.empty((nv, 2 * c), dtype=torch.bfloat16, device='cuda')
    lib.ssr_tf32_split_bf16(x, x2, nv, c, st)
    whi = torch.empty(lib.ssr_conv3d_packed_size(c, 0, co, 2), dtype=torch.float32, device='cuda')
    w16 = torch.empty(lib.ssr_conv3d_packed_size(c, 0, co, 8), dtype=torch.float32, device='cuda')
    lib.ssr_conv3d_pack_weights(w, whi, c, 0, co, 2, st)
    lib.ssr_conv3d_pack_weights(w, w16, c, 0, co, 8, st)
    for with_sums in (False, True):
        y = torch.full((nv, co), float('nan'), dtype=torch.float32, device='cuda')
        sums = torch.full((2 * co,), float('nan'), dtype=torch.float64, device='cuda') if with_sums else None
        lib.ssr_conv3d_fwd_tc_k2n_part(x, c, 0, c, whi, b, y, 1, *d, co, 1, 0, 0, st)
        lib.ssr_conv3d_fwd_tc_k2n_bf16(x2, 2 * c, w16, b, y, sums, 1, *d, co, 1, st)
        torch.cuda.synchronize()
        y64 = _conv64(x, w, b, d)
        err = (y.double().cpu() - y64).abs().max().item() / y64.abs().max().item()
        _log('hybrid k2n %s sums=%d: max/max %.2e' % (d, with_sums, err))
        assert err < KERNEL_TOL, err
        if with_sums:
            s = sums.cpu().numpy()
            assert np.allclose(s[:co], y.double().sum(0).cpu().numpy(), rtol=1e-6, atol=1e-6 * nv)
            assert np.allclose(s[co:], (y.double() ** 2).sum(0).cpu().numpy(), rtol=1e-6, atol=1e-6 * nv)


@pytest.mark.parametrize('dl,cu,co', [([8, 8, 16], 96, 48), ([10, 12, 8], 192, 96), ([8, 8, 8], 48, 24)])
@pytest.mark.parametrize('level', [4, 5])
def test_hybrid_parity_forward_matches_float64(dl, cu, co, level):
    from synthsr_b200._lib import lib, stream_ptr
    split, pm = _SCHEME[level]
    rng = np.random.default_rng(4)
    cs = co
    nl = int(np.prod(dl))
    df = [2 * v for v in dl]
    low = _t(rng.normal(size=(nl, cu)))
    w = _t(rng.normal(size=(3, 3, 3, cs + cu, co)) / np.sqrt(27 * (cs + cu)))
    st = stream_ptr()
    wskip = torch.empty(27 * cs * co, dtype=torch.float32, device='cuda')
    weff = torch.empty(8 * 27 * cu * co, dtype=torch.float32, device='cuda')
    lib.ssr_conv3d_up_weights(w, cs, cu, co, wskip, weff, st)
    n7 = lib.ssr_conv3d_packed_size(cu, cu, co, pm)
    wp8 = torch.empty(8 * n7, dtype=torch.float32, device='cuda')
    for par in range(8):
        lib.ssr_conv3d_pack_weights(weff[par * 27 * cu * co:], wp8[par * n7:], cu, cu, co, pm, st)
    low2 = torch.empty((nl, 2 * cu), dtype=torch.bfloat16, device='cuda')
    getattr(lib, split)(low, low2, nl, cu, st)
    y = torch.full((8 * nl, co), float('nan'), dtype=torch.float32, device='cuda')
    lib.ssr_conv3d_fwd_tc_up_comp(low, low2, cu, wp8, y, 1, *dl, co, level, st)
    torch.cuda.synchronize()
    up = low.view(*dl, cu).repeat_interleave(2, 0).repeat_interleave(2, 1).repeat_interleave(2, 2).reshape(-1, cu)
    y64 = _conv64(up, w[:, :, :, cs:, :], None, df, elu=False)
    err = (y.double().cpu() - y64).abs().max().item() / y64.abs().max().item()
    _log('%s parity %s %d->%d: max/max %.2e' % (split, dl, cu, co, err))
    assert err < KERNEL_TOL, err


def test_producers_emit_the_same_bf16_split_as_the_standalone_pass():
    """ssr_conv3d_first_fwd_split and ssr_conv3d_fwd_tc_k2n_part_split write [bf16(y_lo) | bf16(y_hi)] from their epilogues:
    bit-identical to ssr_tf32_split_bf16 of their output; the whole step agrees with the fusion switched off."""
    from synthsr_b200._lib import lib, stream_ptr
    from synthsr_b200.unet import UNet3D
    rng = np.random.default_rng(5)
    d, co = [12, 20, 36], 24
    nv = int(np.prod(d))
    st = stream_ptr()
    for cin in (1, 2):
        x = _t(rng.normal(size=(nv, cin)))
        w = _t(rng.normal(size=(3, 3, 3, cin, co)) / np.sqrt(27 * cin))
        b = _t(rng.normal(size=co))
        y, y_ref = torch.empty((nv, co), device='cuda'), torch.empty((nv, co), device='cuda')
        y2 = torch.zeros((nv, 2 * co), dtype=torch.bfloat16, device='cuda')
        y2_ref = torch.zeros_like(y2)
        lib.ssr_conv3d_first_fwd_split(x, cin, w, b, y, y2, 1, *d, co, 1, st)
        lib.ssr_conv3d_fwd_ref(x, cin, None, 0, w, b, y_ref, 1, *d, co, 3, 1, st)
        lib.ssr_tf32_split_bf16(y_ref, y2_ref, nv, co, st)
        torch.cuda.synchronize()
        assert torch.equal(y, y_ref) and torch.equal(y2.view(torch.int16), y2_ref.view(torch.int16)), cin
    x = _t(rng.normal(size=(nv, co)))
    w = _t(rng.normal(size=(3, 3, 3, co, co)) / np.sqrt(27 * co))
    b = _t(rng.normal(size=co))
    wp = torch.empty(lib.ssr_conv3d_packed_size(co, 0, co, 2), dtype=torch.float32, device='cuda')
    lib.ssr_conv3d_pack_weights(w, wp, co, 0, co, 2, st)
    part = _t(rng.normal(size=(nv, co)))
    for acc in (0, 1):
        y, y_ref = part.clone(), part.clone()
        y2 = torch.zeros((nv, 2 * co), dtype=torch.bfloat16, device='cuda')
        y2_ref = torch.zeros_like(y2)
        lib.ssr_conv3d_fwd_tc_k2n_part_split(x, co, 0, co, wp, b, y, y2, 1, *d, co, 1, acc, st)
        lib.ssr_conv3d_fwd_tc_k2n_part(x, co, 0, co, wp, b, y_ref, 1, *d, co, 1, acc, 1, st)
        lib.ssr_tf32_split_bf16(y_ref, y2_ref, nv, co, st)
        torch.cuda.synchronize()
        assert torch.equal(y, y_ref) and torch.equal(y2.view(torch.int16), y2_ref.view(torch.int16)), acc
    # whole step: fused vs separate split passes (3 levels: an 8^3 bottleneck, not the chaotic 2^3 one of 5 levels at 32^3)
    image, target = _t(rng.uniform(0, 1, size=(1, 32, 32, 32, 1))), _t(rng.uniform(0, 1, size=(1, 32, 32, 32, 1)))
    out = []
    for fused in (True, False):
        if not fused:
            os.environ['SSR_NO_SPLIT_FUSION'] = '1'
        try:
            net = UNet3D([32, 32, 32, 1], nb_levels=3, batchsize=1, conv_impl='tc3', seed=0)
            loss = net.loss_and_grad(image, target)
            torch.cuda.synchronize()
            out.append((loss.item(), net.pred.clone()))
        finally:
            os.environ.pop('SSR_NO_SPLIT_FUSION', None)
    # (not bit-identical run to run: BatchNorm sums and split-K partial sums are accumulated with atomics in a varying order,
    # and a randomly initialised net amplifies the last-bit differences to 1e-5 .. 1e-4 -- two runs of the SAME configuration
    # differ by as much; the bit-exactness of the fused split itself is asserted above)
    assert abs(out[0][0] - out[1][0]) <= 1e-4 * abs(out[1][0])
    assert (out[0][1] - out[1][1]).abs().max().item() <= 1e-3 * out[1][1].abs().max().item()


# ---------------------------------------------------------------------------------------------------------------------
def _step_vs_routed_oracle(net, image, target, tag, nb_levels=5, **loss_kw):
    """one training step of `net`, then the float64 oracle on the same inputs WITH THE GPU FORWARD'S MAX-POOL WINNERS
    (oracle.unet._maxpool_routed): the only setting in which gradients can be compared tensor by tensor at every size.
    Also asserts that the imposed routing is a max-pool of the float64 forward up to near-ties: the entry it picks is within
    5e-5 (relative to the tensor's range) of the float64 window maximum in every window (measured: <= 5.3e-6; a handful of
    windows differ on noise inputs, ~1 % on inputs with flat background, where the fp32 entries tie exactly)."""
    loss = net.loss_and_grad(torch.from_numpy(image).cuda(), torch.from_numpy(target).cuda(), **loss_kw)
    torch.cuda.synchronize()
    routing = _gpu_routing(net)
    report = {}
    pred_o, loss_o, grads_o = _oracle64(net.state_dict(), image, target, nb_levels, routing=routing, report=report, **loss_kw)
    for l, (nflip, nwin, gap) in report.items():
        _log('%s: level %d pooling, %d of %d windows routed differently from float64 argmax, largest gap %.1e' % (tag, l, nflip, nwin, gap))
        assert gap <= 5e-5, (l, nflip, nwin, gap)
    return _errors_of(net, loss, pred_o, loss_o, grads_o, tag + ' (same pooling winners)')


def _step_errors(net, image, target, pred_ref, loss_ref, grads_ref, tag, **loss_kw):
    loss = net.loss_and_grad(torch.from_numpy(image).cuda(), torch.from_numpy(target).cuda(), **loss_kw)
    torch.cuda.synchronize()
    return _errors_of(net, loss, pred_ref, loss_ref, grads_ref, tag)


def _errors_of(net, loss, pred_ref, loss_ref, grads_ref, tag):
    pred = net.pred.view(pred_ref.shape).cpu().numpy().astype(np.float64)
    e_l2 = np.linalg.norm(pred - pred_ref) / np.linalg.norm(pred_ref)
    e_max = np.abs(pred - pred_ref).max() / np.abs(pred_ref).max()
    e_loss = abs(loss.item() - loss_ref) / abs(loss_ref)
    # per tensor, relative to max(|tensor|, 1e-2 |whole gradient|) (tensors that are analytically ~0 are rounding noise)
    gtot = np.sqrt(sum(float((np.asarray(g, np.float64) ** 2).sum()) for g in grads_ref.values()))
    gerr = {k: np.linalg.norm(net.g[k].cpu().numpy().astype(np.float64) - np.asarray(g, np.float64)) /
            max(np.linalg.norm(np.asarray(g, np.float64)), 1e-2 * gtot) for k, g in grads_ref.items()}
    worst = max(gerr, key=gerr.get)
    gall = np.sqrt(sum(float(((net.g[k].cpu().numpy().astype(np.float64) - np.asarray(g, np.float64)) ** 2).sum())
                       for k, g in grads_ref.items())) / gtot
    gerr = dict(gerr)
    gerr['__whole_gradient__'] = gall
    _log('%s: pred relL2 %.3e max/max %.3e loss rel %.3e whole-gradient relL2 %.3e worst tensor %.3e (%s)' % (
        tag, e_l2, e_max, e_loss, gall, gerr[worst], worst))
    return e_l2, e_max, e_loss, gerr


def _assert_north_star(e_l2, e_max, e_loss, gerr):
    assert e_l2 <= PRED_TOL, ('prediction relative L2', e_l2)
    assert e_max <= PRED_TOL, ('prediction max/max', e_max)
    assert e_loss <= LOSS_TOL, ('loss', e_loss)
    for k, e in gerr.items():
        assert e <= GRAD_TOL, ('gradient', k, e)


def _oracle64(sd, image, target, nb_levels, routing=None, report=None, **loss_kw):
    from oracle import unet as OU
    params = {k: torch.tensor(np.asarray(v), dtype=torch.float64) for k, v in sd.items()}
    names = OU.trainable_names(params)
    leaves = {k: params[k].clone().requires_grad_(True) for k in names}
    p = {k: leaves.get(k, params[k]) for k in params}
    img, tgt = torch.tensor(image, dtype=torch.float64), torch.tensor(target, dtype=torch.float64)
    pred = OU.forward(p, img, training=True, nb_levels=nb_levels, pool_routing=routing, routing_report=report)
    loss = OU.loss_fn(pred, img, tgt, **loss_kw)
    grads = dict(zip(names, torch.autograd.grad(loss, [leaves[k] for k in names])))
    return pred.detach().numpy(), float(loss.detach()), {k: v.numpy() for k, v in grads.items()}


def _generated_batch(size, seed=0):
    """one batch the way the benchmark produces it: bench.py's label phantom through the CUDA generator with training()'s
    default hyper-parameters (SynthSR/training.py:57-73) -> (image, target) float32 [1, n, n, n, 1]"""
    import bench
    from synthsr_b200.draws import sample_draws
    from synthsr_b200.generator import GeneratorPlan, SynthGenerator
    maps, pm, ps, gl, gc = bench.make_inputs(size, 1, seed=seed)
    plan = GeneratorPlan([size] * 3, True, 0, gl, None, 1., None, **bench.TRAINING_DEFAULTS)
    rng = np.random.default_rng(seed)
    gen = SynthGenerator(plan, 1)
    m, sd = bench.draw_gmm(rng, pm, ps, gc)
    image, target = gen.run(torch.from_numpy(maps[0][None]).cuda(), m, sd, sample_draws(rng, plan, 1))
    torch.cuda.synchronize()
    return image.cpu().numpy().copy(), target.cpu().numpy().copy()


def _noise(size, seed=1):
    rng = np.random.default_rng(seed)
    return (rng.uniform(0, 1, size=(1, size, size, size, 1)).astype(np.float32),
            rng.uniform(0, 1, size=(1, size, size, size, 1)).astype(np.float32))


def _both_modes_vs_oracle(size, image, target, tag, metric='l1', state=None):
    """exact-fp32 mode and 'tc3' against the float64 oracle on the same step -> (errors of ref, errors of tc3)"""
    from synthsr_b200.unet import UNet3D
    dims = [size] * 3
    out = {}
    for impl in ('ref', 'tc3'):
        net = UNet3D(dims + [1], batchsize=1, conv_impl=impl, seed=0)
        if state is not None:
            net.load_state_dict(state)
        if impl == 'ref':
            pred_o, loss_o, grads_o = _oracle64(net.state_dict(), image, target, 5, metric=metric)
        out[impl] = _step_errors(net, image, target, pred_o, loss_o, grads_o, '%s %s' % (impl, tag), metric=metric)
        del net
        torch.cuda.empty_cache()
    return out['ref'], out['tc3']


@pytest.mark.parametrize('size,metric', [(32, 'l1'), (32, 'l2'), (64, 'l1'), (64, 'l2')])
def test_tc3_training_step_meets_north_star_vs_float64_oracle(size, metric):
    """random init (glorot, seed 0), uniform-noise image and target -- the hardest input for error amplification.  Every bar,
    every tensor (float64 oracle evaluated with the GPU forward's max-pool winners, see _step_vs_routed_oracle)."""
    from synthsr_b200.unet import UNet3D
    dims = [size] * 3
    net = UNet3D(dims + [1], batchsize=1, conv_impl='tc3', seed=0)
    image, target = _noise(size)
    errs = _step_vs_routed_oracle(net, image, target, 'tc3 %d^3 %s random init, noise, vs float64 oracle' % (size, metric),
                                  metric=metric)
    _assert_north_star(*errs)


@pytest.mark.parametrize('kind,metric', [('noise', 'l2'), ('noise', 'l1'), ('generated', 'l1')])
def test_tc3_training_step_96_vs_float64_oracle(kind, metric):
    """96^3, random init, against the float64 oracle: uniform noise (l2, l1) and a batch of the benchmark's own distribution
    (label phantom -> CUDA generator).  Every bar, every tensor, with the oracle's pooling routed like the GPU forward's
    (without that, ONE near-tied window decides a first-level tensor's error: see the next test)."""
    from synthsr_b200.unet import UNet3D
    image, target = _noise(96) if kind == 'noise' else _generated_batch(96)
    net = UNet3D([96, 96, 96, 1], batchsize=1, conv_impl='tc3', seed=0)
    errs = _step_vs_routed_oracle(net, image, target, 'tc3 96^3 %s random init, %s, vs float64 oracle' % (metric, kind),
                                  metric=metric)
    _assert_north_star(*errs)


def test_gradient_comparison_is_limited_by_maxpool_argmax_flips():
    """Why the tests above impose the GPU forward's pooling winners on the oracle.  MaxPooling3D routes each window's
    gradient to its argmax, and two forwards that differ in the last bits disagree on the argmax of a few near-tied windows.
    ONE flipped window among N_w moves that level's activation gradient by sqrt(2 / N_w) in relative L2 -- 3.5e-3 for the
    166k windows of level 2 at 96^3 -- and every shallower weight gradient inherits it (scripts/actgrad_diag.py,
    profiles/r02_actgrad_96_noise_l2.txt: dL/dpre of the EXACT-fp32 mode is 1e-5 from float64 down to level 3 and jumps to
    4.9e-3 at the level-2 pooling; at 48^3 the same jump is 7e-5, at 32^3 it does not occur).  No fp32 implementation --
    the reference's own TF fp32 included -- can be held to a per-tensor bar against float64 under free routing.
    Pinned here with the exact-fp32 mode at 96^3, noise inputs: against the free-running oracle some tensor is > 1e-3 off
    while the whole gradient stays far inside 1e-2; against the oracle routed like its own forward every tensor is within
    1e-3 (the same step, the same numbers -- only the handful of near-tied windows resolved the same way)."""
    from synthsr_b200.unet import UNet3D
    image, target = _noise(96)
    ref, tc3 = _both_modes_vs_oracle(96, image, target, '96^3 l2 random init, noise (argmax-flip study, free routing)', 'l2')
    worst_ref = max(v for k, v in ref[3].items() if k != '__whole_gradient__')
    assert worst_ref > WELL_POSED, worst_ref
    assert ref[3]['__whole_gradient__'] <= GRAD_TOL and tc3[3]['__whole_gradient__'] <= GRAD_TOL
    net = UNet3D([96, 96, 96, 1], batchsize=1, conv_impl='ref', seed=0)
    routed = _step_vs_routed_oracle(net, image, target, 'ref 96^3 l2 random init, noise (argmax-flip study)', metric='l2')
    assert max(routed[3].values()) <= WELL_POSED, max(routed[3].values())


def test_tc3_training_step_trained_reference_weights_real_scan():
    """the reference's trained weights (models/SynthSR_v10_210712.h5) on a 96^3 crop of data/images/brain1.nii.gz: every
    bar, every tensor."""
    wfile, image = _find('models/SynthSR_v10_210712.h5'), _find('images/brain1.nii.gz')
    if wfile is None or image is None:
        pytest.skip('reference weights / scan not available on this machine')
    from SynthSR import predict as P
    from ext.lab2im import utils
    from synthsr_b200 import h5lite
    from synthsr_b200.unet import UNet3D
    im, aff, _ = utils.load_volume(image, im_only=False, dtype='float')
    S = P.preprocess(im, aff)[0]
    c = [s // 2 - 48 for s in S.shape[1:4]]
    crop = np.ascontiguousarray(S[:, c[0]:c[0] + 96, c[1]:c[1] + 96, c[2]:c[2] + 96, :], dtype=np.float32)
    target = np.ascontiguousarray(np.roll(crop, 1, axis=1) * .5 + crop * .5)       # a smooth, image-like target
    sd, _ = h5lite.load_keras_weights(wfile)
    net = UNet3D([96, 96, 96, 1], batchsize=1, conv_impl='tc3', seed=0)
    net.load_state_dict(sd)
    errs = _step_vs_routed_oracle(net, crop, target, 'tc3 96^3 trained reference weights, brain1 crop')
    _assert_north_star(*errs)


def test_tc3_training_step_at_benchmark_size_160():
    """BASELINE configs[1]: 160^3, batch 1, reference topology, random init, a batch generated exactly as bench.py does (and
    uniform noise).  Reference = the exact-fp32 CUDA-core mode on the same device (a float64 CPU step at 160^3 needs ~30 GB;
    the mode is validated against the float64 oracle at the sizes the CPU reaches).  Prediction, loss and the whole gradient
    at north_star's bars; individual tensors are recorded (at this size the exact-fp32 mode cannot serve as a per-tensor
    reference: it is itself 5e-3 off float64 at 96^3, see the argmax-flip test).  The plain-TF32 fast mode is recorded next to
    it, not gated."""
    from synthsr_b200.unet import UNet3D
    dims = [160] * 3
    cases = {'generated batch': _generated_batch(160), 'noise': _noise(160)}
    refs = {}
    ref = UNet3D(dims + [1], batchsize=1, conv_impl='ref', seed=0)
    for tag, (image, target) in cases.items():
        loss_r = ref.loss_and_grad(torch.from_numpy(image).cuda(), torch.from_numpy(target).cuda())
        torch.cuda.synchronize()
        refs[tag] = (ref.pred.view(1, *dims, 1).cpu().numpy().astype(np.float64), loss_r.item(),
                     {k: ref.g[k].cpu().numpy().astype(np.float64) for k in ref.layout})
    del ref
    torch.cuda.empty_cache()
    for impl in ('tc3', 'tc'):
        net = UNet3D(dims + [1], batchsize=1, conv_impl=impl, seed=0)
        for tag, (image, target) in cases.items():
            e_l2, e_max, e_loss, gerr = _step_errors(net, image, target, *refs[tag],
                                                     '%s 160^3 l1 random init, %s, vs exact-fp32 mode' % (impl, tag))
            if impl == 'tc3':
                assert e_l2 <= PRED_TOL and e_max <= PRED_TOL and e_loss <= LOSS_TOL, (tag, e_l2, e_max, e_loss)
                assert gerr['__whole_gradient__'] <= GRAD_TOL, (tag, gerr['__whole_gradient__'])
        del net
        torch.cuda.empty_cache()
