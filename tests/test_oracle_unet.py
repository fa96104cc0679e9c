"""CPU tests pinning the torch U-Net oracle: topology / parameter count of the reference model, Keras layer names,
analytic known-answer cases, and the Keras Adam / BN-moving-average formulas."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import unet as OU


def test_parameter_count_and_names():
    """13,242,049 parameters for Cin=1 (13,237,633 conv + 4,416 BN), Keras layer names (3.4)."""
    specs = OU.layer_specs(1)
    n_conv = sum(27 * ci * co + co for n, k, ci, co in specs if k == 'conv') + sum(ci * co + co for n, k, ci, co in specs if k == 'conv1')
    n_bn = sum(4 * co for n, k, ci, co in specs if k == "bn")   # gamma, beta + non-trainable moving mean/variance
    assert n_conv == 13237633 and n_bn == 4416 and n_conv + n_bn == 13242049
    names = [n for n, *_ in specs]
    assert names[0] == 'unet_conv_downarm_0_0' and 'unet_bn_down_4' in names and 'unet_conv_uparm_8_1' in names
    assert 'unet_bn_up_3' in names and names[-1] == 'unet_likelihood'
    cins = {n: ci for n, k, ci, co in specs}
    assert cins['unet_conv_uparm_5_0'] == 576 and cins['unet_conv_uparm_8_0'] == 72      # concat widths (8a-U2)
    from synthsr_b200.unet import layer_specs
    assert layer_specs(1) == specs and layer_specs(2, 8, 3) == OU.layer_specs(2, 8, 3)


def test_zero_weights_output_is_bias():
    p = OU.init_params(0, 1, nb_features=4, nb_levels=2)
    for k in p:
        if k.endswith('kernel'):
            p[k].zero_()
    p['unet_likelihood/bias'].fill_(0.25)
    x = torch.rand(1, 8, 8, 8, 1)
    y = OU.forward(p, x, training=True, nb_levels=2)
    assert y.shape == (1, 8, 8, 8, 1) and torch.allclose(y, torch.full_like(y, 0.25))


def test_skip_is_pre_bn_conv_output():
    """the skip tensor is the second conv's post-ELU, pre-BN output (models.py:431-432): scaling BN gamma of level 0
    must not change the skip half of the decoder input."""
    torch.manual_seed(0)
    p = OU.init_params(1, 1, nb_features=4, nb_levels=2)
    x = torch.rand(1, 8, 8, 8, 1)
    acts = {}
    OU.forward(p, x, nb_levels=2, activations=acts)
    skip = acts['unet_conv_downarm_0_1'].clone()
    p['unet_bn_down_0/gamma'] *= 3.
    acts2 = {}
    OU.forward(p, x, nb_levels=2, activations=acts2)
    assert torch.equal(acts2['unet_conv_downarm_0_1'], skip)


def test_maxpool_same_odd():
    x = torch.arange(27, dtype=torch.float32).reshape(1, 1, 3, 3, 3)
    y = OU._maxpool_same(x)
    assert y.shape == (1, 1, 2, 2, 2) and y[0, 0, 1, 1, 1] == 26 and y[0, 0, 0, 0, 0] == 13


def test_routed_maxpool_is_the_free_one_under_its_own_winners():
    """pool_routing (test aid of the GPU gradient-parity tests): with the forward's own window winners the routed pooling is
    MaxPooling3D -- same prediction, same gradients; with one near-tied window resolved the other way the prediction
    barely moves and the report says so"""
    import torch.nn.functional as F
    params = OU.init_params(0, 1, dtype=torch.float64, nb_features=4, nb_levels=3)
    leaves = {k: params[k].clone().requires_grad_(True) for k in OU.trainable_names(params)}
    p = {k: leaves.get(k, params[k]) for k in params}
    img = torch.rand(1, 8, 8, 8, 1, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    x, routing = img.permute(0, 4, 1, 2, 3), []
    for level in range(2):
        for j in range(2):
            x = F.elu(OU._conv(x, params, 'unet_conv_downarm_%d_%d' % (level, j)))
        x = OU._bn(x, params, 'unet_bn_down_%d' % level, True, None)
        x, idx = F.max_pool3d(x, 2, return_indices=True)
        routing.append(idx)
    free = OU.forward(p, img, nb_levels=3)
    g_free = torch.autograd.grad(free.square().sum(), list(leaves.values()))
    report = {}
    routed = OU.forward(p, img, nb_levels=3, pool_routing=routing, routing_report=report)
    g_routed = torch.autograd.grad(routed.square().sum(), list(leaves.values()))
    assert torch.equal(free, routed) and all(torch.equal(a, b) for a, b in zip(g_free, g_routed))
    assert report == {0: (0, routing[0].numel(), 0.), 1: (0, routing[1].numel(), 0.)}
    other = [routing[0].clone(), routing[1]]
    other[0][0, 0, 0, 0, 0] = 0 if routing[0][0, 0, 0, 0, 0] != 0 else 1       # another entry of the first window
    report = {}
    OU.forward(p, img, nb_levels=3, pool_routing=other, routing_report=report)
    assert report[0][0] == 1 and report[0][2] > 0 and report[1][0] == 0


def test_adam_and_bn_moving_formulas():
    p = OU.init_params(2, 1, nb_features=4, nb_levels=2)
    p = {k: v.double() for k, v in p.items()}
    opt = OU.adam_init(p)
    x, t = torch.rand(1, 8, 8, 8, 1, dtype=torch.float64), torch.rand(1, 8, 8, 8, 1, dtype=torch.float64)
    p0 = {k: v.clone() for k, v in p.items()}
    # forward() default nb_levels=5 does not fit this tiny net: use the generic step below
    names = OU.trainable_names(p)
    leaves = {k: p[k].clone().requires_grad_(True) for k in names}
    q = {k: leaves.get(k, p[k]) for k in p}
    new = {}
    acts = {}
    pred = OU.forward(q, x, nb_levels=2, new_stats=new, activations=acts)
    loss = OU.loss_fn(pred, x, t)
    g = torch.autograd.grad(loss, [leaves[k] for k in names])
    # first Adam step: m = .1 g, v = .001 g^2, lr_t = lr*sqrt(.001)/.1 -> update = lr * g/(|g| + eps*...) ~ lr*sign(g)
    lr = 1e-3
    lr_t = lr * math.sqrt(1 - .999) / (1 - .9)
    k0 = names[0]
    upd = lr_t * (.1 * g[0]) / ((.001 * g[0] ** 2).sqrt() + 1e-7)
    assert torch.allclose(upd.abs().max(), torch.tensor(lr, dtype=torch.float64), rtol=1e-3)
    # BN moving variance uses n/(n-(1+eps)) (Keras 2.3.1)
    h = acts['unet_conv_downarm_0_1']
    n = h.numel() / h.shape[1]
    var = h.var(dim=(0, 2, 3, 4), unbiased=False)
    exp = .99 * 1. + .01 * var * n / (n - (1 + 1e-3))
    assert torch.allclose(new['unet_bn_down_0/moving_variance'], exp.detach())


def test_train_step_reduces_loss():
    torch.manual_seed(0)
    p = OU.init_params(3, 1)
    opt = OU.adam_init(p)
    x, t = torch.rand(1, 16, 16, 16, 1), torch.rand(1, 16, 16, 16, 1)
    l0, _, _ = OU.train_step(p, opt, x, t, lr=1e-3)
    for _ in range(3):
        l1, _, _ = OU.train_step(p, opt, x, t, lr=1e-3)
    assert l1 < l0 and opt['iterations'] == 4


@pytest.mark.skipif(not os.path.isfile('/root/reference/models/SynthSR_v10_210712.h5'),
                    reason='reference weights / scan are only mounted in the build container')
def test_trained_reference_weights_pin_the_layer_semantics():
    """TensorFlow / Keras cannot run here, so the oracle's reading of the Keras graph (U3: cross-correlation kernels in
    (kd,kh,kw,Cin,Cout) layout, ELU, skip taken before BatchNorm, concat order [skip, upsampled], inference BatchNorm on
    the moving statistics) has no TF output to be compared with.  The reference ships something almost as good: a model
    TRAINED under those semantics (models/SynthSR_v10_210712.h5) and a 1 mm scan it was meant for
    (data/images/brain1.nii.gz).  Under the oracle's reading the network reproduces the anatomy of its input (correlation
    > 0.85 on a 96^3 crop); under every plausible misreading it falls apart."""
    import torch
    from oracle import unet as OU
    from SynthSR import predict as P
    from ext.lab2im import utils
    from synthsr_b200 import h5lite
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    sd, _ = h5lite.load_keras_weights('/root/reference/models/SynthSR_v10_210712.h5')
    params = {k: torch.tensor(v) for k, v in sd.items()}
    im, aff, _ = utils.load_volume('/root/reference/data/images/brain1.nii.gz', im_only=False, dtype='float')
    S, idx, shape, _ = P.preprocess(im, aff)
    c = [s // 2 - 48 for s in S.shape[1:4]]
    crop = np.ascontiguousarray(S[:, c[0]:c[0] + 96, c[1]:c[1] + 96, c[2]:c[2] + 96, :], dtype=np.float32)
    x = torch.from_numpy(crop)

    def corr(p, **kw):
        with torch.no_grad():
            pred = OU.forward(p, x, training=kw.pop('training', False), **kw).numpy()
        a, b = pred[0, 16:-16, 16:-16, 16:-16, 0].ravel(), crop[0, 16:-16, 16:-16, 16:-16, 0].ravel()
        return float(np.corrcoef(a, b)[0, 1])

    right = corr(params)
    flipped = {k: (v.flip(0, 1, 2) if k.endswith('kernel') and v.shape[0] == 3 else v) for k, v in params.items()}
    transposed = {k: (v.permute(2, 1, 0, 3, 4).contiguous() if k.endswith('kernel') and v.shape[0] == 3 else v)
                  for k, v in params.items()}
    wrong = {'true convolution (flipped kernels)': corr(flipped),
             'kernel axes in (kw,kh,kd) order': corr(transposed),
             'concat order [upsampled, skip]': corr(params, _wrong=('concat_swapped',)),
             'skip taken after BatchNorm': corr(params, _wrong=('skip_after_bn',)),
             'ReLU instead of ELU': corr(params, _wrong=('relu',)),
             'batch statistics at inference': corr(params, training=True)}
    assert right > 0.85, right
    for name, r in wrong.items():
        assert r < right - 0.15, (name, r, right)


def test_loss_matches_reference_metrics_model():
    """oracle loss_fn vs the reference's own metrics_model() executed on the tf shim (tests/golden/make_reference_loss_goldens.py):
    residual addition from 'image_out' channels, centre cropping (including odd margins), L1 / L2.  Tolerance 2e-6 relative:
    only the order of the mean's summation differs."""
    import os
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_loss.npz'))
    cases = {'l1_plain': dict(metric='l1'),
             'l2_crop_res': dict(metric='l2', loss_cropping=8, work_with_residual_channel=[0]),
             'l1_crop3_res2': dict(metric='l1', loss_cropping=[8, 10, 6], work_with_residual_channel=[0, 2]),
             'l1_crop_odd': dict(metric='l1', loss_cropping=6)}
    for name, kw in cases.items():
        loss = OU.loss_fn(torch.from_numpy(G[name + '_pred']), torch.from_numpy(G[name + '_image']),
                          torch.from_numpy(G[name + '_target']), **kw)
        np.testing.assert_allclose(float(loss), float(G[name + '_loss']), rtol=2e-6)


def test_graph_built_by_the_reference_unet_function():
    """ext/neuron/models.unet() (-> conv_enc, conv_dec) executed unmodified on a functional-API stand-in for Keras
    (tests/golden/make_reference_unet_goldens.py): the layers it creates -- names, order, kernel shapes -- are exactly the
    product's layer_specs / keras_layer_order, and the oracle's forward on the same weights reproduces the prediction and the
    intermediate activations of THAT graph (skip taken from conv_downarm_l_1's output, concat [skip, up], BN after the second
    activation, linear 1x1x1 head)."""
    import os
    from synthsr_b200.unet import keras_layer_order, layer_specs
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_unet.npz'))
    assert [str(n) for n in G['layer_order']] == keras_layer_order(nb_levels=3)
    weights = {k[2:]: G[k] for k in G.files if k.startswith('w/')}
    specs = layer_specs(2, nb_features=4, nb_levels=3, feat_mult=2, nb_conv_per_level=2, nb_labels=1)
    want = {}
    for name, kind, ci, co in specs:
        if kind == 'bn':
            want[name + '/gamma'], want[name + '/beta'] = (co,), (co,)
        else:
            ks = 3 if kind == 'conv' else 1
            want[name + '/kernel'], want[name + '/bias'] = (ks, ks, ks, ci, co), (co,)
    assert {k: tuple(v.shape) for k, v in weights.items()} == want
    params = OU.init_params(0, 2, dtype=torch.float64, nb_features=4, nb_levels=3, feat_mult=2, nb_conv_per_level=2, nb_labels=1)
    for k, v in weights.items():
        assert tuple(params[k].shape) == tuple(v.shape), k
        params[k] = torch.from_numpy(v.astype(np.float64))
    acts = {}
    pred = OU.forward(params, torch.from_numpy(G['image'].astype(np.float64)), training=True, nb_levels=3, activations=acts)
    np.testing.assert_allclose(pred.numpy(), G['prediction'], rtol=0, atol=1e-10)
    np.testing.assert_allclose(acts['unet_conv_downarm_0_1'].permute(0, 2, 3, 4, 1).numpy(), G['act/unet_conv_downarm_0_1'],
                               rtol=0, atol=1e-12)
    # the wrong readings really are different graphs
    for wrong in ('concat_swapped', 'skip_after_bn', 'relu'):
        p2 = OU.forward(params, torch.from_numpy(G['image'].astype(np.float64)), training=True, nb_levels=3, _wrong=(wrong,))
        assert np.abs(p2.numpy() - G['prediction']).max() > 1e-2, wrong


def test_segmentation_regularised_loss_matches_reference():
    """SURVEY.md 8f rank 4, oracle only (no GPU side yet): the reference's own add_seg_loss_to_model() + DiceLoss executed on
    the stand-ins, its frozen segmentation network built by the reference's unet(final_pred_activation='softmax')
    (tests/golden/make_reference_segloss_goldens.py), against oracle seg_regularised_loss -- label merging through the
    equivalency table, ignored labels, centre cropping, intensity clipping, the FreeSurfer-header axis swap."""
    import os
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_segloss.npz'))
    cases = {'plain': dict(rel_weight=.25), 'crop_clip_fs': dict(rel_weight=.5, loss_cropping=8, m=.1, M=.8, fs_header=True)}
    for name, kw in cases.items():
        pre = name + '_w/'
        weights = {k[len(pre):]: G[k] for k in G.files if k.startswith(pre)}
        n_seg = len(G[name + '_equiv'])
        params = OU.init_params(0, 1, dtype=torch.float64, nb_features=4, nb_levels=3, feat_mult=2, nb_conv_per_level=2,
                                nb_labels=n_seg)
        assert {k for k in params if not k.endswith(('moving_mean', 'moving_variance'))} == set(weights)
        for k, v in weights.items():
            params[k] = torch.from_numpy(v.astype(np.float64))
        total = OU.seg_regularised_loss(float(G[name + '_image_loss']), torch.from_numpy(G[name + '_pred_image'].astype(np.float64)),
                                        torch.from_numpy(G[name + '_seg_target']), params, G['generation_labels'],
                                        G[name + '_equiv'], nb_levels=3, **kw)
        np.testing.assert_allclose(float(total), float(G[name + '_total']), rtol=1e-9)


def test_class_tables_formulation_of_the_dice_matches_reference():
    """the (cls_of_seg, gt_value) tables the CUDA Dice kernels take (synthsr_b200.seg_loss.class_tables) express the reference's
    label-merging loop: softmax -> merge by table -> soft Dice on the reference-executed golden gives the reference's total."""
    import os
    from synthsr_b200.seg_loss import class_tables
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_segloss.npz'))
    name = 'plain'
    pre = name + '_w/'
    weights = {k[len(pre):]: G[k] for k in G.files if k.startswith(pre)}
    equiv = G[name + '_equiv']
    params = OU.init_params(0, 1, dtype=torch.float64, nb_features=4, nb_levels=3, feat_mult=2, nb_conv_per_level=2,
                            nb_labels=len(equiv))
    for k, v in weights.items():
        params[k] = torch.from_numpy(v.astype(np.float64))
    logits = OU.forward(params, torch.from_numpy(G[name + '_pred_image'].astype(np.float64)), training=True, nb_levels=3)
    cls, gtv = class_tables(G['generation_labels'], equiv)
    s = torch.softmax(logits, -1)
    K = len(gtv)
    p = torch.stack([sum(s[..., j] for j in range(len(cls)) if cls[j] == k) for k in range(K)], -1)
    lab = torch.from_numpy(G[name + '_seg_target'])[..., 0]
    gt = torch.stack([(lab == int(gtv[k])).double() for k in range(K)], -1)
    top, bot = (2 * gt * p).sum((1, 2, 3)), (gt ** 2 + p ** 2).sum((1, 2, 3))
    total = float(G[name + '_image_loss']) + .25 * float((1 - (top + 1e-7) / (bot + 1e-7)).mean())
    np.testing.assert_allclose(total, float(G[name + '_total']), rtol=1e-9)
