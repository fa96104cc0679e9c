"""Per-entry-point cross-check of the generator kernels against tests/host_emulator.py (the NumPy stand-in the CPU suite uses
for the same C-ABI calls): every call is issued twice with identical arguments -- once to libsynthsr_b200.so on CUDA tensors,
once to the emulator on host copies -- and the outputs compared.  Finer-grained than tests/test_generator_gpu.py (which
checks whole graphs), meant for localising a failure.

Runs by default (first validated on a B200 in round 2, gpurun_out call A: 6 passed); SSR_KERNEL_CROSSCHECK=0 skips it."""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get('SSR_KERNEL_CROSSCHECK') == '0', reason='SSR_KERNEL_CROSSCHECK=0')]

f32 = np.float32


@pytest.fixture
def both():
    from host_emulator import HostEmulator
    from synthsr_b200._lib import lib, stream_ptr
    return lib, HostEmulator(), stream_ptr()


def _pair(a):
    h = torch.from_numpy(np.ascontiguousarray(a))
    return h.cuda(), h.clone()


def _same(g, h, atol):
    torch.cuda.synchronize()
    np.testing.assert_allclose(g.cpu().numpy(), h.numpy(), rtol=0, atol=atol)


def test_resize_linear_nearest_strided(both):
    lib, emu, st = both
    rng = np.random.default_rng(1)
    for C, nearest, (s, d) in [(1, 0, ([9, 7, 11], [16, 16, 12])), (3, 0, ([5, 5, 5], [20, 24, 16])), (1, 1, ([16, 12, 20], [16, 12, 6]))]:
        xg, xh = _pair(rng.normal(size=(2, *s, C)).astype(f32))
        for stride, off in ((0, 0), (4, 2)) if C == 1 else ((0, 0),):
            og, oh = _pair(np.full((2, *d, max(stride, C)), 7., f32))
            lib.ssr_resize(xg, og, 2, *s, *d, C, nearest, stride, off, st)
            emu.ssr_resize(xh, oh, 2, *s, *d, C, nearest, stride, off, 0)
            _same(og, oh, 1e-6)


def test_svf_integrate(both):
    lib, emu, st = both
    rng = np.random.default_rng(2)
    vg, vh = _pair((rng.normal(size=(1, 10, 12, 8, 3)) * 2).astype(f32))
    tg = torch.empty_like(vg)
    lib.ssr_svf_integrate(vg, tg, 1, 10, 12, 8, 7, st)
    emu.ssr_svf_integrate(vh, None, 1, 10, 12, 8, 7, 0)
    _same(vg, vh, 0)                                          # same float32 op order, no FMA: bit exact


def test_blur3d_plain_and_fused_normalisation(both):
    lib, emu, st = both
    rng = np.random.default_rng(3)
    n = [12, 10, 14]
    xg, xh = _pair(rng.uniform(0, 300, size=(2, *n)).astype(f32))
    for ksh in ([3, 3, 3], [3, 3, 5], [1, 1, 7]):
        k = rng.uniform(size=ksh).astype(f32)
        k /= k.sum()
        kg, kh = _pair(k)
        og, oh = _pair(np.zeros((2, *n, 2), f32))
        lib.ssr_blur3d(xg, og, kg, *ksh, None, None, 2, *n, 1, 0, 2, 1, st)
        emu.ssr_blur3d(xh, oh, kh, *ksh, None, None, 2, *n, 1, 0, 2, 1, 0)
        _same(og, oh, 1e-3)
    # fused min-max normalisation + gamma: produce the min/max with the library's / the emulator's own ssr_minmax
    mg, mh = torch.empty((2, 2), dtype=torch.int32, device='cuda'), torch.empty((2, 2), dtype=torch.int32)
    lib.ssr_minmax(xg, mg, 2, int(np.prod(n)), st)
    emu.ssr_minmax(xh, mh, 2, int(np.prod(n)), 0)
    gg, gh = _pair(np.array([.8, 1.3], f32))
    kg, kh = _pair(np.full((3, 3, 3), 1 / 27., f32))
    og, oh = _pair(np.zeros((2, *n), f32))
    lib.ssr_blur3d(xg, og, kg, 3, 3, 3, mg, gg, 2, *n, 1, 0, 1, 0, st)
    emu.ssr_blur3d(xh, oh, kh, 3, 3, 3, mh, gh, 2, *n, 1, 0, 1, 0, 0)
    _same(og, oh, 1e-5)


def test_warp_linear_affine_field_crop_flip(both):
    lib, emu, st = both
    from synthsr_b200 import draws as D
    rng = np.random.default_rng(4)
    n, p, h, c = [20, 16, 18], [2, 0, 1], [10, 8, 9], [14, 12, 12]
    inner = [n[i] - 2 * p[i] for i in range(3)]
    xg, xh = _pair(rng.uniform(size=(2, *inner)).astype(f32))
    aff = np.stack([D.build_affine(rng.uniform(-10, 10, 3).astype(f32), None, rng.uniform(.9, 1.1, 3).astype(f32),
                                   rng.uniform(-2, 2, 3).astype(f32)) for _ in range(2)])
    ag, ah = _pair(aff)
    fg, fh = _pair(rng.normal(size=(2, *h, 3)).astype(f32))
    cg, ch = _pair(np.array([[3, 2, 4], [0, 4, 6]], np.int32))
    flg, flh = _pair(np.array([1, 0], np.uint8))
    og, oh = _pair(np.zeros((2, *c), f32))
    lib.ssr_warp_linear(xg, og, ag, fg, 2, *n, *p, *h, cg, *c, flg, st)
    emu.ssr_warp_linear(xh, oh, ah, fh, 2, *n, *p, *h, ch, *c, flh, 0)
    _same(og, oh, 1e-5)


def test_deform_labels_bit_exact(both):
    lib, emu, st = both
    from synthsr_b200 import draws as D
    rng = np.random.default_rng(5)
    n, p, h, c = [20, 16, 18], [0, 2, 1], [10, 8, 9], [16, 12, 14]
    inner = [n[i] - 2 * p[i] for i in range(3)]
    lg, lh = _pair(rng.integers(0, 9, size=(2, *inner)).astype(np.int32))
    aff = np.stack([D.build_affine(rng.uniform(-15, 15, 3).astype(f32), rng.uniform(-.01, .01, 6).astype(f32),
                                   rng.uniform(.85, 1.15, 3).astype(f32), None) for _ in range(2)])
    ag, ah = _pair(aff)
    fg, fh = _pair((rng.normal(size=(2, *h, 3)) * 1.5).astype(f32))
    cg, ch = _pair(np.array([[3, 2, 4], [0, 4, 0]], np.int32))
    flg, flh = _pair(np.array([0, 1], np.uint8))
    lut = np.arange(9, dtype=np.int32)
    lut[[3, 4, 5, 6]] = [5, 6, 3, 4]
    tg, th = _pair(lut)
    og, oh = _pair(np.zeros((2, *c), np.int32))
    lib.ssr_deform_labels_nearest(lg, og, ag, fg, 2, *n, *p, *h, cg, *c, flg, tg, 9, st)
    emu.ssr_deform_labels_nearest(lh, oh, ah, fh, 2, *n, *p, *h, ch, *c, flh, th, 9, 0)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(og.cpu().numpy(), oh.numpy())


def test_mimic_acquisition_and_reliability_fill(both):
    lib, emu, st = both
    from synthsr_b200.generator import mimic_zooms, reliability_factors
    rng = np.random.default_rng(6)
    n, o = [16, 14, 18], [16, 14, 18]
    xg, xh = _pair(rng.uniform(size=(2, *n)).astype(f32))
    params = np.stack([mimic_zooms(n, [1., 1., 1.], r, o) for r in ([1., 1., 1.], [2.3, 1., 6.7])])
    pg, ph = _pair(params)
    og, oh = _pair(np.zeros((2, *o, 2), f32))
    lib.ssr_mimic_acquisition(xg, og, og, pg, 2, *n, *o, 2, 0, 2, 1, st)
    emu.ssr_mimic_acquisition(xh, oh, oh, ph, 2, *n, *o, 2, 0, 2, 1, 0)
    _same(og, oh, 1e-5)
    fs = reliability_factors(o, [16, 14, 6])
    fg = [torch.from_numpy(f).cuda() for f in fs]
    fh = [torch.from_numpy(f.copy()) for f in fs]
    lib.ssr_fill_outer3(og, *fg, 2, *o, 2, 1, st)
    emu.ssr_fill_outer3(oh, *fh, 2, *o, 2, 1, 0)
    _same(og, oh, 0)
