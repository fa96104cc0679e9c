"""Segmentation-regularised loss on the GPU (synthsr_b200/seg_loss.py, csrc/seg_loss.cu) against
float64 torch references and the oracle (oracle/unet.py:seg_regularised_loss, pinned by executing the reference).

First run on a B200 in round 2 (kernels passed as written; the step test runs in the default 'tc3' mode at north_star's bars).
SSR_ENABLE_SEG_LOSS=0 skips the file."""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get('SSR_ENABLE_SEG_LOSS') == '0', reason='SSR_ENABLE_SEG_LOSS=0')]


def _dice_ref(logits, labels, cls, gtv, rel_weight, crop):
    """float64 torch: softmax -> merge -> soft Dice (DiceLoss, enable_checks=False) -> rel_weight * mean(1 - dice)."""
    s = torch.softmax(logits, -1)                                            # [B,X,Y,Z,S]
    K = len(gtv)
    p = torch.stack([sum(s[..., j] for j in range(len(cls)) if cls[j] == k) for k in range(K)], -1)
    gt = torch.stack([(labels == int(gtv[k])).double() for k in range(K)], -1)
    if crop is not None:
        (c, b) = crop
        sl = (slice(None),) + tuple(slice(b[i], b[i] + c[i]) for i in range(3))
        p, gt = p[sl], gt[sl]
    top = (2 * gt * p).sum((1, 2, 3))
    bot = (gt ** 2 + p ** 2).sum((1, 2, 3))
    return rel_weight * (1 - (top + 1e-7) / (bot + 1e-7)).mean()


@pytest.mark.parametrize('crop', [None, ([8, 6, 10], [2, 3, 1])])
def test_softmax_dice_kernels(crop):
    import ctypes
    from synthsr_b200._lib import lib, stream_ptr
    rng = np.random.default_rng(3)
    B, d, S = 2, [12, 12, 14], 6
    cls = np.array([0, 1, 1, -1, 2, 0], np.int32)
    gtv = np.array([0, 2, 5], np.int32)
    logits = rng.normal(size=(B, *d, S)).astype(np.float32) * 2
    labels = rng.integers(0, 7, size=(B, *d)).astype(np.int32)
    lg = torch.from_numpy(logits).cuda()
    lab = torch.from_numpy(labels).cuda()
    cg, gg = torch.from_numpy(cls).cuda(), torch.from_numpy(gtv).cuda()
    sums = torch.empty(B * 3 * 2, dtype=torch.float64, device='cuda')
    loss = torch.full((1,), .125, dtype=torch.float64, device='cuda')
    dl = torch.full((B, *d, S), float('nan'), device='cuda')
    cs = cb = None
    if crop is not None:
        keep = ((ctypes.c_int * 3)(*crop[0]), (ctypes.c_int * 3)(*crop[1]))
        cs, cb = ctypes.cast(keep[0], ctypes.c_void_p), ctypes.cast(keep[1], ctypes.c_void_p)
    st = stream_ptr()
    lib.ssr_softmax_dice_sums(lg, S, lab, cg, gg, 3, B, *d, cs, cb, sums, st)
    lib.ssr_dice_finalize(sums, B, 3, .25, loss, st)
    lib.ssr_softmax_dice_grad(lg, S, lab, cg, gg, 3, B, *d, cs, cb, sums, .25, dl, st)
    torch.cuda.synchronize()
    z = torch.from_numpy(logits).double().requires_grad_(True)
    ref = _dice_ref(z, torch.from_numpy(labels), cls, gtv, .25, crop)
    g, = torch.autograd.grad(ref, z)
    np.testing.assert_allclose(loss.item() - .125, float(ref.detach()), rtol=1e-5)
    assert torch.isfinite(dl).all()
    scale = g.abs().max().item()
    np.testing.assert_allclose(dl.cpu().double().numpy(), g.numpy(), rtol=0, atol=2e-5 * scale)


def test_seg_input_and_head_extra_grad():
    from synthsr_b200._lib import lib, stream_ptr
    rng = np.random.default_rng(4)
    V, C = 3000, 24
    pred = torch.from_numpy(rng.uniform(-.3, 1.3, size=V).astype(np.float32)).cuda()
    image = torch.from_numpy(rng.uniform(-.2, .2, size=(V, 3)).astype(np.float32)).cuda()
    dy = torch.from_numpy(rng.normal(size=V).astype(np.float32)).cuda()
    st = stream_ptr()
    for img, clip in ((None, 0), (image, 1)):
        y, dp = torch.empty(V, device='cuda'), torch.empty(V, device='cuda')
        lib.ssr_seg_input(pred, img, 3 if img is not None else 0, 1, clip, .1, .9, y, V, st)
        lib.ssr_seg_input_bwd(pred, img, 3 if img is not None else 0, 1, clip, .1, .9, dy, dp, V, st)
        torch.cuda.synchronize()
        x = pred.double() + (image[:, 1].double() if img is not None else 0)
        x = x.clone().requires_grad_(True)
        ref = (torch.clamp(x, .1, .9) - .1) / .8 if clip else x + 0.
        g, = torch.autograd.grad((ref * dy.double()).sum(), x)
        np.testing.assert_allclose(y.cpu().numpy(), ref.detach().cpu().numpy(), rtol=0, atol=1e-6)
        np.testing.assert_allclose(dp.cpu().numpy(), g.cpu().numpy(), rtol=0, atol=1e-6)
    feat = torch.from_numpy(rng.normal(size=(V, C)).astype(np.float32)).cuda()
    w = torch.from_numpy(rng.normal(size=C).astype(np.float32)).cuda()
    e = torch.from_numpy(rng.normal(size=V).astype(np.float32)).cuda()
    stats = torch.from_numpy(rng.uniform(.5, 1.5, size=4 * C).astype(np.float32)).cuda()
    for s in (None, stats):
        dfeat0 = torch.from_numpy(rng.normal(size=(V, C)).astype(np.float32)).cuda()
        dfeat, dw, db = dfeat0.clone(), torch.full((C,), 2., device='cuda'), torch.full((1,), -1., device='cuda')
        lib.ssr_head_extra_grad(feat, s, w, e, V, C, dfeat, dw, db, st)
        torch.cuda.synchronize()
        f = feat.double() * (s[2 * C:3 * C].double() if s is not None else 1.) + (s[3 * C:].double() if s is not None else 0.)
        np.testing.assert_allclose(dfeat.cpu().numpy(), (dfeat0.double() + e.double()[:, None] * w.double()[None]).cpu().numpy(),
                                   rtol=0, atol=1e-5)
        np.testing.assert_allclose(dw.cpu().numpy() - 2., (f * e.double()[:, None]).sum(0).cpu().numpy(), rtol=1e-4, atol=1e-3)
        np.testing.assert_allclose(db.item() + 1., e.double().sum().item(), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize('impl,tol', [('ref', 2e-4), ('tc3', 1e-3)])
def test_step_with_segmentation_regulariser_matches_oracle(impl, tol):
    """one training step of a small network (3 levels, 8 features) with a small frozen segmentation network attached: loss
    and every gradient of the trained network against float64 autograd through the oracle's seg_regularised_loss."""
    from oracle import unet as OU
    from synthsr_b200.seg_loss import SegRegularisedUNet3D, SegRegulariser, class_tables
    from synthsr_b200.unet import UNet3D
    rng = np.random.default_rng(7)
    dims, B, L, F, S = [16, 16, 16], 1, 3, 8, 5
    gen_labels = np.array([0, 1, 2, 3, 4, 14, 15])
    equiv = np.array([0, 2, 2, 3, -1])
    tmp = UNet3D(dims + [1], nb_features=F, nb_levels=L, nb_labels=S, batchsize=B, conv_impl='ref', seed=5)
    seg_sd = tmp.state_dict()
    for k in seg_sd:                                          # non-trivial BN parameters and biases
        if k.endswith(('gamma',)):
            seg_sd[k] = rng.uniform(.7, 1.3, size=seg_sd[k].shape).astype(np.float32)
        elif k.endswith(('beta', 'bias')):
            seg_sd[k] = (rng.normal(size=seg_sd[k].shape) * .1).astype(np.float32)
    del tmp
    seg = SegRegulariser(dims, B, seg_sd, S, gen_labels, equiv, rel_weight=.5, loss_cropping=12, m=.05, M=.9,
                         nb_features=F, nb_levels=L, conv_impl=impl)
    net = SegRegularisedUNet3D(dims + [2], nb_features=F, nb_levels=L, nb_labels=1, batchsize=B, conv_impl=impl, seed=1, seg=seg)
    image = rng.uniform(0, 1, size=(B, *dims, 2)).astype(np.float32)
    target = rng.uniform(0, 1, size=(B, *dims, 1)).astype(np.float32)
    labels = gen_labels[rng.integers(0, len(gen_labels), size=(B, *dims))].astype(np.int32)
    labels[:, :8] = rng.integers(0, 5, size=(B, 8, 16, 16))    # values that equal loop indices, so the Dice has ground truth
    net.seg_labels = torch.from_numpy(labels).cuda()
    loss = net.loss_and_grad(torch.from_numpy(image).cuda(), torch.from_numpy(target).cuda(), 'l1', [0], 12)
    torch.cuda.synchronize()
    # ---- oracle, float64 autograd
    params = {k: torch.tensor(v, dtype=torch.float64) for k, v in net.state_dict().items()}
    names = OU.trainable_names(params)
    for k in names:
        params[k].requires_grad_(True)
    img_t, tgt_t = torch.from_numpy(image).double(), torch.from_numpy(target).double()
    pred = OU.forward(params, img_t, training=True, nb_levels=L)
    image_loss = OU.loss_fn(pred, img_t, tgt_t, metric='l1', work_with_residual_channel=[0], loss_cropping=12)
    predicted_image = pred + img_t[..., 0:1]
    seg_params = {k: torch.tensor(v, dtype=torch.float64) for k, v in seg_sd.items()}
    total = OU.seg_regularised_loss(image_loss, predicted_image, torch.from_numpy(labels)[..., None], seg_params, gen_labels,
                                    equiv, .5, loss_cropping=12, m=.05, M=.9, nb_levels=L)
    grads = torch.autograd.grad(total, [params[k] for k in names])
    assert abs(loss.item() - float(total.detach())) <= tol * abs(float(total.detach())), (loss.item(), float(total.detach()))
    assert float(total.detach()) > float(image_loss.detach()) + .05           # the Dice term really contributes
    gtot = np.sqrt(sum(float((g ** 2).sum()) for g in grads))
    for k, g in zip(names, grads):
        err = np.linalg.norm(net.g[k].cpu().numpy().astype(np.float64) - g.numpy()) / max(np.linalg.norm(g.numpy()), 1e-2 * gtot)
        assert err < (5e-4 if impl == 'ref' else 1e-2), (k, err)          # north_star bars in the default mode
    assert class_tables(gen_labels, equiv)[1].tolist() == [0, 2, 3]
