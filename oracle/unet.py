"""torch-CPU restatement of the reference's U-Net + loss + Adam training step.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- PARITY UNPINNED.

Topology follows ext/neuron/models.py:256-498 as called from SynthSR/training.py:330-341
(unet(24, shape, 5 levels, 3x3x3, nb_labels, feat_mult=2, nb_conv_per_level=2, batch_norm=-1, 'elu', linear head));
the loss follows SynthSR/metrics_model.py:53-104; the optimiser follows Keras 2.3.1 `Adam` as used at
SynthSR/training.py:444.  Keras/TF layer arithmetic is third-party (not under /root/reference) and restated from
the published definitions: Conv3D = cross-correlation + bias, 'same' zero padding; ELU(alpha=1);
BatchNormalization(axis=-1, eps=1e-3, momentum=.99) using biased batch variance in training and updating the
moving variance with var*n/(n-(1+eps)); MaxPooling3D(2,'same'); UpSampling3D(2) nearest; Adam with
lr_t = lr*sqrt(1-b2^t)/(1-b1^t), p -= lr_t*m/(sqrt(v)+1e-7).

Data layout at the API is the reference's: activations [B,X,Y,Z,C], kernels (k,k,k,Cin,Cout); internally torch NCDHW.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3
BN_MOMENTUM = 0.99


def layer_specs(cin, nb_features=24, nb_levels=5, feat_mult=2, nb_conv_per_level=2, nb_labels=1, conv_size=3):
    """Ordered list of (name, kind, cin, cout) with the Keras layer names (ext/neuron/models.py:314,350,442,476,480)."""
    specs = []
    c = cin
    enc_feats = []
    for level in range(nb_levels):
        f = int(np.round(nb_features * feat_mult ** level))                      # models.py:306
        for j in range(nb_conv_per_level):
            specs.append(('unet_conv_downarm_%d_%d' % (level, j), 'conv', c, f))
            c = f
        specs.append(('unet_bn_down_%d' % level, 'bn', f, f))
        enc_feats.append(f)
    for level in range(nb_levels - 1):
        f = int(np.round(nb_features * feat_mult ** (nb_levels - 2 - level)))    # models.py:421
        c = enc_feats[nb_levels - 2 - level] + c                                 # concat [skip, up] models.py:434
        for j in range(nb_conv_per_level):
            specs.append(('unet_conv_uparm_%d_%d' % (nb_levels + level, j), 'conv', c, f))
            c = f
        specs.append(('unet_bn_up_%d' % level, 'bn', f, f))
    specs.append(('unet_likelihood', 'conv1', c, nb_labels))                     # models.py:480-481
    return specs


def init_params(seed, cin, dtype=torch.float32, conv_size=3, **kw):
    """glorot_uniform kernels, zero biases, BN gamma=1 beta=0 moving_mean=0 moving_var=1 (Keras defaults)."""
    rng = np.random.default_rng(seed)
    params = OrderedDict()
    for name, kind, ci, co in layer_specs(cin, conv_size=conv_size, **kw):
        if kind in ('conv', 'conv1'):
            k = conv_size if kind == 'conv' else 1
            limit = math.sqrt(6.0 / (k ** 3 * ci + k ** 3 * co))
            w = rng.uniform(-limit, limit, size=(k, k, k, ci, co)).astype(np.float32)
            params[name + '/kernel'] = torch.tensor(w, dtype=dtype)
            params[name + '/bias'] = torch.zeros(co, dtype=dtype)
        else:
            params[name + '/gamma'] = torch.ones(co, dtype=dtype)
            params[name + '/beta'] = torch.zeros(co, dtype=dtype)
            params[name + '/moving_mean'] = torch.zeros(co, dtype=dtype)
            params[name + '/moving_variance'] = torch.ones(co, dtype=dtype)
    return params


def trainable_names(params):
    return [k for k in params if not k.endswith(('moving_mean', 'moving_variance'))]


def _conv(x, params, name):
    w = params[name + '/kernel'].permute(4, 3, 0, 1, 2)          # (k,k,k,Cin,Cout) -> (Cout,Cin,k,k,k)
    k = w.shape[-1]
    return F.conv3d(x, w, params[name + '/bias'], padding=k // 2)


def _bn(x, params, name, training, new_stats):
    g, b = params[name + '/gamma'], params[name + '/beta']
    if training:
        mean = x.mean(dim=(0, 2, 3, 4))
        var = x.var(dim=(0, 2, 3, 4), unbiased=False)
        n = x.numel() / x.shape[1]
        if new_stats is not None:
            mm, mv = params[name + '/moving_mean'], params[name + '/moving_variance']
            new_stats[name + '/moving_mean'] = (mm * BN_MOMENTUM + mean.detach() * (1 - BN_MOMENTUM))
            new_stats[name + '/moving_variance'] = (mv * BN_MOMENTUM +
                                                    var.detach() * (n / (n - (1.0 + BN_EPS))) * (1 - BN_MOMENTUM))
    else:
        mean, var = params[name + '/moving_mean'], params[name + '/moving_variance']
    sh = (1, -1, 1, 1, 1)
    return (x - mean.view(sh)) * torch.rsqrt(var.view(sh) + BN_EPS) * g.view(sh) + b.view(sh)


def _maxpool_same(x):
    """MaxPooling3D(2, padding='same'): pad at the end with -inf for odd sizes."""
    pad = []
    for d in (4, 3, 2):
        pad += [0, x.shape[d] % 2]
    if any(pad):
        x = F.pad(x, pad, value=float('-inf'))
    return F.max_pool3d(x, 2)


def _maxpool_routed(x, idx, report, level):
    """MaxPooling3D with the window winners GIVEN (idx: what F.max_pool3d(..., return_indices=True) returned for another
    implementation's forward of the same tensor).  Comparing two implementations' GRADIENTS tensor by tensor is only
    well-posed under the same routing: a window whose two largest entries agree to the last bits is won by a different
    entry in two forwards that differ by rounding, and one such window changes the whole level's gradient (see
    tests/test_unet_parity_gpu.py).  report[level] = (windows routed differently from this forward's own argmax,
    windows, largest (own max - routed entry) / max|x|): how far the given routing is from this forward's max-pool."""
    assert all(x.shape[d] % 2 == 0 for d in (2, 3, 4)), 'routed pooling: even sizes only'
    flat = x.flatten(2)
    out = flat.gather(2, idx.flatten(2)).view(idx.shape)
    if report is not None:
        own, own_idx = F.max_pool3d(x.detach(), 2, return_indices=True)
        report[level] = (int((own_idx != idx).sum()), idx.numel(),
                         float((own - out.detach()).max() / x.detach().abs().max()))
    return out


def forward(params, image, training=True, nb_levels=5, nb_conv_per_level=2, new_stats=None, activations=None,
            _wrong=(), pool_routing=None, routing_report=None):
    """image [B,X,Y,Z,Cin] -> prediction [B,X,Y,Z,nb_labels].  (ext/neuron/models.py:301-360, 420-498)
    pool_routing (test aid, see _maxpool_routed): per encoder level, the max-pool winners to use instead of this forward's own.
    _wrong: deliberately wrong readings of the graph ('concat_swapped', 'skip_after_bn', 'relu'), only for the test that
    shows the reference's trained weights reject them (tests/test_oracle_unet.py)."""
    act = F.relu if 'relu' in _wrong else F.elu
    x = image.permute(0, 4, 1, 2, 3)
    skips = []
    for level in range(nb_levels):
        for j in range(nb_conv_per_level):
            x = act(_conv(x, params, 'unet_conv_downarm_%d_%d' % (level, j)))
            if activations is not None:
                activations['unet_conv_downarm_%d_%d' % (level, j)] = x
        if 'skip_after_bn' not in _wrong:
            skips.append(x)                                        # conv output, pre-BN (models.py:431-432)
        x = _bn(x, params, 'unet_bn_down_%d' % level, training, new_stats)
        if 'skip_after_bn' in _wrong:
            skips.append(x)
        if level < nb_levels - 1:
            x = _maxpool_same(x) if pool_routing is None else _maxpool_routed(x, pool_routing[level], routing_report, level)
    for level in range(nb_levels - 1):
        x = F.interpolate(x, scale_factor=2, mode='nearest')       # UpSampling3D (models.py:425-427)
        pair = [skips[nb_levels - 2 - level], x]
        x = torch.cat(pair[::-1] if 'concat_swapped' in _wrong else pair, dim=1)    # models.py:434
        for j in range(nb_conv_per_level):
            x = act(_conv(x, params, 'unet_conv_uparm_%d_%d' % (nb_levels + level, j)))
            if activations is not None:
                activations['unet_conv_uparm_%d_%d' % (nb_levels + level, j)] = x
        x = _bn(x, params, 'unet_bn_up_%d' % level, training, new_stats)
    x = _conv(x, params, 'unet_likelihood')
    return x.permute(0, 2, 3, 4, 1)


def loss_fn(pred, image, target, metric='l1', work_with_residual_channel=None, loss_cropping=None):
    """SynthSR/metrics_model.py:53-104."""
    if work_with_residual_channel is not None:
        res = torch.stack([image[..., c] for c in work_with_residual_channel], -1)   # :57-62 (image_out channels)
        pred = res + pred                                                            # :65
    if loss_cropping is not None:
        shp = target.shape[1:4]
        lc = [loss_cropping] * 3 if isinstance(loss_cropping, int) else list(loss_cropping)
        b = [int((shp[i] - lc[i]) / 2) for i in range(3)]                            # :79
        sl = (slice(None),) + tuple(slice(b[i], b[i] + lc[i]) for i in range(3))
        pred, target = pred[sl], target[sl]
    err = pred - target
    if metric == 'l1':
        return err.abs().mean()                                                      # :104
    if metric == 'l2':
        return (err ** 2).mean()                                                     # :101
    raise NotImplementedError(metric)


def seg_regularised_loss(image_loss, predicted_image, seg_target, seg_params, generation_labels, equivalency, rel_weight,
                         loss_cropping=None, m=None, M=None, fs_header=False, nb_levels=5):
    """SynthSR/metrics_model.py:136-215 (`add_seg_loss_to_model`) + ext/lab2im/layers.py:1334-1376 (`DiceLoss`,
    enable_checks=False, no class / boundary weights): image loss + rel_weight * soft Dice between the one-hot segmentation
    target and the prediction of a frozen segmentation U-Net (softmax head) applied to the predicted image.
    SURVEY.md 8f rank 4 -- oracle only, the GPU side is not built yet.  predicted_image [B,X,Y,Z,1], seg_target [B,X,Y,Z,1] int.

    Kept as the reference has them: the ground-truth map of generation label i is `seg_target == i` -- the loop INDEX, not the
    label value generation_labels[i] (:188); the frozen network's BatchNorm normalises with batch statistics while fitting
    (Keras 2.3.1 BatchNormalization.call ignores `trainable`; restated, unpinned)."""
    x = predicted_image
    if m is not None:
        x = (torch.clamp(x, m, M) - m) / (M - m)                                       # :154
    if fs_header:
        x = x.permute(0, 1, 3, 2, 4).flip(2)                                           # :158
    seg = torch.softmax(forward(seg_params, x, training=True, nb_levels=nb_levels), dim=-1)
    if fs_header:
        seg = seg.flip(2).permute(0, 1, 3, 2, 4)                                       # :161-162
    tgt = seg_target
    if loss_cropping is not None:                                                      # :167-183
        shp = predicted_image.shape[1:4]
        lc = [loss_cropping] * 3 if isinstance(loss_cropping, int) else list(loss_cropping)
        b = [int((shp[i] - lc[i]) / 2) for i in range(3)]
        sl = (slice(None),) + tuple(slice(b[i], b[i] + lc[i]) for i in range(3))
        tgt, seg = tgt[sl], seg[sl]
    equivalency = np.asarray(equivalency)
    gts, preds = [], []
    for i in range(len(generation_labels)):                                            # :189-204
        idx = np.where(equivalency == generation_labels[i])[0]
        if len(idx) > 0:
            if len(idx) > 3:
                raise Exception("uuummm weird that you're merging so many labels...")
            gts.append((tgt[..., -1] == i).to(seg.dtype))
            preds.append(sum(seg[..., int(j)] for j in idx))
    gt, pred = torch.stack(gts, -1), torch.stack(preds, -1)
    top = (2 * gt * pred).sum(dim=(1, 2, 3))                                           # layers.py:1344, 1361
    bottom = (gt ** 2 + pred ** 2).sum(dim=(1, 2, 3))
    dice = (top + 1e-7) / (bottom + 1e-7)
    return image_loss + rel_weight * (1 - dice).mean()                                 # layers.py:1364, 1376; :209


def adam_init(params):
    return {'iterations': 0, 'm': {k: torch.zeros_like(params[k]) for k in trainable_names(params)},
            'v': {k: torch.zeros_like(params[k]) for k in trainable_names(params)}}


def train_step(params, opt, image, target, lr=1e-4, lr_decay=0., beta1=.9, beta2=.999, eps=1e-7, **loss_kw):
    """One Keras train_on_batch: forward (training BN), L1 loss, backward, Adam.  Updates params/opt in place.
    Returns (loss, grads dict, prediction)."""
    names = trainable_names(params)
    leaves = {k: params[k].detach().clone().requires_grad_(True) for k in names}
    p = OrderedDict((k, leaves.get(k, params[k])) for k in params)
    new_stats = {}
    pred = forward(p, image, training=True, new_stats=new_stats)
    loss = loss_fn(pred, image, target, **loss_kw)
    grads = torch.autograd.grad(loss, [leaves[k] for k in names])
    grads = dict(zip(names, grads))
    lr_eff = lr
    if lr_decay > 0:
        lr_eff = lr * (1. / (1. + lr_decay * opt['iterations']))
    t = opt['iterations'] + 1
    lr_t = lr_eff * (math.sqrt(1. - beta2 ** t) / (1. - beta1 ** t))
    with torch.no_grad():
        for k in names:
            g = grads[k]
            opt['m'][k] = beta1 * opt['m'][k] + (1. - beta1) * g
            opt['v'][k] = beta2 * opt['v'][k] + (1. - beta2) * g * g
            params[k] = params[k] - lr_t * opt['m'][k] / (opt['v'][k].sqrt() + eps)
        for k, v in new_stats.items():
            params[k] = v
    opt['iterations'] = t
    return float(loss.detach()), grads, pred.detach()
