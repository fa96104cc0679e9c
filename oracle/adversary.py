"""CPU oracle of the adversarial fine-tuner (TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU legs
may import it): a float64 restatement of SynthSR/fine_tuning_with_adversary.py's discriminator and losses, written without
the product's building blocks (no F.conv3d / F.pad: convolutions are 27 shifted-slice matrix products), differentiable twice
through torch autograd.

PINNED by the reference's own code executed on the tf shim (tests/golden/make_reference_adversary_goldens.py ->
reference_adversary.npz / reference_adversary_wiring.json): build_generator_loss, build_discriminator_loss (with the
gradient array given), RandomWeightedAverage, the layer sequence of make_discriminator.
UNPINNED (restated from the Keras / TensorFlow documentation; TensorFlow cannot run in this image): the arithmetic of
Conv3D(padding='same', strides=2) -- TensorFlow pads (ceil(n/s) - 1) * s + k - n voxels with the odd one at the END --,
Dense, Flatten(channels_last), LeakyReLU, and K.gradients (= d sum(outputs) / d input).
"""
import numpy as np
import torch


def tf_same_padding(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def conv3d_same(x, w, b, stride):
    """KL.Conv3D(filters, 3, strides=stride, padding='same') on x [B, X, Y, Z, Cin]; w (3, 3, 3, Cin, Cout) (Keras layout)."""
    B, dims, k = x.shape[0], list(x.shape[1:4]), w.shape[0]
    pads = [tf_same_padding(n, k, stride) for n in dims]
    xp = x.new_zeros((B, *[n + p[0] + p[1] for n, p in zip(dims, pads)], x.shape[-1]))
    core = tuple(slice(p[0], p[0] + n) for n, p in zip(dims, pads))
    xp = xp.clone()
    xp[(slice(None),) + core] = x
    out = [-(-n // stride) for n in dims]
    y = 0.
    for k0 in range(k):
        for k1 in range(k):
            for k2 in range(k):
                patch = xp[:, k0:k0 + stride * out[0]:stride, k1:k1 + stride * out[1]:stride, k2:k2 + stride * out[2]:stride, :]
                y = y + patch @ w[k0, k1, k2]
    return y + b


def leaky_relu(x, alpha=0.2):
    return torch.where(x > 0, x, alpha * x)


def discriminator_forward(params, x, mask=None, n_levels=4):
    """fine_tuning_with_adversary.py:482-508.  params: {'conv3d_<i>/kernel' (3,3,3,Cin,Cout), '.../bias', 'dense_1/...',
    'dense_2/...'} in layer order; x [B, X, Y, Z, C] -> [B, 1]."""
    t = x if mask is None else x * mask.to(x.dtype)                                   # :486-487
    i = 0
    for level in range(n_levels):                                                     # :491-493
        for stride in (1, 2):
            i += 1
            t = leaky_relu(conv3d_same(t, params['conv3d_%d/kernel' % i], params['conv3d_%d/bias' % i], stride))   # :511-514
    t = t.reshape(t.shape[0], -1)                                                     # Flatten(channels_last), :495
    t = leaky_relu(t @ params['dense_1/kernel'] + params['dense_1/bias'])             # :496-497
    return t @ params['dense_2/kernel'] + params['dense_2/bias']                      # :500


def random_weighted_average(real, fake, weights):
    """:619-624; weights [B, 1, 1, 1, 1]"""
    return weights * real + (1 - weights) * fake


def gradient_penalty_terms(d_real, d_fake, gradients, gradient_penalty_w=10, n_dims=3):
    """build_discriminator_loss (:580-596) given the gradients of the discriminator at the averaged samples."""
    norm = torch.sqrt((gradients ** 2).sum(dim=tuple(range(1, n_dims + 1))))         # :585, spatial axes only
    penalty = gradient_penalty_w * (1 - norm) ** 2                                    # :586
    return (-d_real).mean() + d_fake.mean() + penalty.mean()                          # :589-594


def discriminator_loss(params, real, fake, weights, gradient_penalty_w=10, mask=None, n_levels=4):
    avg = random_weighted_average(real, fake, weights).detach().requires_grad_(True)
    d_real = discriminator_forward(params, real, mask, n_levels)
    d_fake = discriminator_forward(params, fake, mask, n_levels)
    d_avg = discriminator_forward(params, avg, mask, n_levels)
    grads = torch.autograd.grad(d_avg.sum(), avg, create_graph=True)[0]               # Gradients layer, :640
    return gradient_penalty_terms(d_real, d_fake, grads, gradient_penalty_w, real.dim() - 2)


def generator_loss(target, pred, d_out, discr_weight, loss_cropping=None, target_seg=None, seg_out=None,
                   generation_labels=None, segmentation_equivalency=None, dice_weight=0.):
    """build_generator_loss (:511-577): l1_weight * L1 + discr_weight * mean(-D) [+ dice_weight * Dice], l1_weight = 1 -
    discr_weight [- dice_weight].  The ground truth of a class is `target_seg == ll`, the label VALUE (:551)."""
    use_seg = seg_out is not None
    if loss_cropping is not None:                                                     # :516-538
        shp = list(target.shape[1:-1])
        lc = [int(loss_cropping)] * len(shp) if np.isscalar(loss_cropping) else [int(v) for v in loss_cropping]
        b = [int((shp[i] - lc[i]) / 2) for i in range(len(shp))]
        sl = (slice(None),) + tuple(slice(b[i], b[i] + lc[i]) for i in range(len(shp)))
        target, pred = target[sl], pred[sl]
        if use_seg:
            target_seg, seg_out = target_seg[sl], seg_out[sl]
    l1 = (target - pred).abs().mean()                                                 # :541
    w = (-d_out).mean()                                                               # :542
    l1_weight = 1 - discr_weight                                                      # :568
    if not use_seg:
        return l1_weight * l1 + discr_weight * w                                      # :574-575
    gts, preds = [], []
    eq = np.asarray(segmentation_equivalency)
    for ll in generation_labels:                                                      # :548-561
        idx = np.where(eq == ll)[0]
        if len(idx) > 0:
            if len(idx) > 3:
                raise Exception("uuummm weird that you're merging so many labels...")
            gts.append((target_seg[..., -1] == int(ll)).to(seg_out.dtype))
            preds.append(sum(seg_out[..., int(j)] for j in idx))
    gt, pr = torch.stack(gts, -1), torch.stack(preds, -1)
    top = (2 * gt * pr).sum(dim=(1, 2, 3))                                            # ext/lab2im/layers.py:1344, 1361
    bottom = (gt ** 2 + pr ** 2).sum(dim=(1, 2, 3))
    dice = (1 - (top + 1e-7) / (bottom + 1e-7)).mean()                                # layers.py:1364, 1376
    return (l1_weight - dice_weight) * l1 + discr_weight * w + dice_weight * dice     # :570-573
