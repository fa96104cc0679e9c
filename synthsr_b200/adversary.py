"""Adversarial (WGAN-GP) fine-tuning of the super-resolution U-Net: the engine side of the reference's
SynthSR/fine_tuning_with_adversary.py (second half).

    Discriminator           `make_discriminator` (fine_tuning_with_adversary.py:482-508): n_levels x [Conv3D(3, stride 1) +
                            LeakyReLU(0.2), Conv3D(3, stride 2) + LeakyReLU(0.2)], Flatten, Dense + LeakyReLU, Dense(1).
    discriminator_loss      `build_discriminator_loss` (:580-596) with `RandomWeightedAverage` (:606-624) and `Gradients` (:627-641).
    AdversarialUNet3D       the network being fine-tuned: its head step turns the image loss into `build_generator_loss` (:511-577):
                            (1 - w_d [- w_s]) * L1 + w_d * mean(-D(prediction)) [+ w_s * Dice].
    AdversarialEngine       the two alternating steps of the reference's loop (:440-459).

What runs where.  The generator, the U-Net (forward, backward, Adam) and the optional segmentation regulariser are this
package's CUDA kernels, exactly as in SynthSR.training.  The DISCRIMINATOR is evaluated with torch autograd (cuDNN
convolutions): its gradient penalty differentiates the norm of the discriminator's input gradient with respect to the
discriminator's weights -- second derivatives through every layer -- and none of the hand-written backward kernels has a
double backward.  Its parameters live in ONE flat buffer so that the same fused Keras-Adam kernel (ssr_adam_flat) updates
them.  The adversarial term reaches the U-Net through the head's extra-gradient entry point (ssr_head_extra_grad), like the
Dice term of the segmentation regulariser.

Keras semantics restated here (TensorFlow cannot run in this image; oracle/adversary.py restates them independently in
float64 and tests/golden/make_reference_adversary_goldens.py executes the reference's own loss builders and layer wiring):
`padding='same'` of a strided convolution pads (total = (ceil(n / s) - 1) * s + k - n) with the extra voxel at the END;
Flatten is channels-last; glorot-uniform kernels and zero biases; a non-trainable network inside a fitted model still
normalises with batch statistics and only drops its moving-average updates.
"""
import math
import os

import numpy as np
import torch
import torch.nn.functional as F

from ._lib import lib, stream_ptr
from .seg_loss import SegRegularisedUNet3D
from .unet import UNet3D

LEAKY_ALPHA = 0.2


def same_padding(n, k, s):
    """TensorFlow's `padding='same'`: (before, after) for an axis of n voxels, kernel k, stride s."""
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def fp32_convs():
    """context for everything that evaluates or differentiates the discriminator on the GPU: the reference convolves in fp32,
    cuDNN's TF32 default costs 4e-4 on the discriminator loss (the flag is read when a kernel runs, backward passes included)"""
    # cuDNN times its algorithms for the discriminator's few shapes (still fp32): 1616 -> 700 ms per discriminator step at
    # 160^3 (scripts/adv_step_time.py); SSR_ADV_CUDNN_BENCHMARK=0 switches that off.  SSR_ADV_TF32=1 additionally lets
    # cuDNN use TF32 (325 ms; outside the fp32 parity the tests check).
    return torch.backends.cudnn.flags(enabled=True, allow_tf32=os.environ.get('SSR_ADV_TF32') == '1',
                                      benchmark=os.environ.get('SSR_ADV_CUDNN_BENCHMARK') != '0')


class Discriminator:
    """fine_tuning_with_adversary.py:482-508.  Tensors are channels-last [B, X, Y, Z, C] at the interface (the layout of
    the engine and of Keras); kernels are kept in the Keras layout (3, 3, 3, Cin, Cout) / (in, out) as views of one flat
    parameter buffer."""

    def __init__(self, input_shape, n_filters=32, n_levels=4, mask_input=False, device='cuda', seed=0, dtype=torch.float32):
        self.input_shape = [int(s) for s in input_shape]
        self.n_filters, self.n_levels, self.mask_input = int(n_filters), int(n_levels), bool(mask_input)
        self.device, self.dtype = torch.device(device), dtype
        dims, cin = self.input_shape[:-1], self.input_shape[-1]
        self.layers = []                                     # (name, kind, stride)
        shapes = []
        for level in range(self.n_levels):
            f = self.n_filters * 2 ** level
            for stride in (1, 2):
                name = 'conv3d_%d' % (len(self.layers) + 1)
                self.layers.append((name, 'conv', stride))
                shapes += [(name + '/kernel', (3, 3, 3, cin, f)), (name + '/bias', (f,))]
                cin = f
                dims = [-(-d // stride) for d in dims]
        flat = int(np.prod(dims)) * cin
        units = self.n_filters * 2 ** self.n_levels
        self.layers += [('dense_1', 'dense', 0), ('dense_2', 'dense', 0)]
        shapes += [('dense_1/kernel', (flat, units)), ('dense_1/bias', (units,)), ('dense_2/kernel', (units, 1)),
                   ('dense_2/bias', (1,))]
        self.layout, n = {}, 0
        for name, shp in shapes:
            self.layout[name] = (n, shp)
            n += (int(np.prod(shp)) + 3) // 4 * 4            # 16-byte aligned tensors
        self.n_params = n
        self.params = torch.zeros(n, dtype=dtype, device=self.device)
        self.grads = torch.zeros(n, dtype=dtype, device=self.device)
        self.adam_m = torch.zeros(n, dtype=dtype, device=self.device)
        self.adam_v = torch.zeros(n, dtype=dtype, device=self.device)
        self.iterations = 0
        self.p = {k: self.params[o:o + int(np.prod(shp))].view(shp) for k, (o, shp) in self.layout.items()}
        self.g = {k: self.grads[o:o + int(np.prod(shp))].view(shp) for k, (o, shp) in self.layout.items()}
        rng = np.random.default_rng(seed)
        for k, (o, shp) in self.layout.items():              # Keras defaults: glorot_uniform kernels, zero biases
            if k.endswith('kernel'):
                rf = int(np.prod(shp[:-2]))
                lim = math.sqrt(6. / (rf * shp[-2] + rf * shp[-1]))
                self.p[k].copy_(torch.as_tensor(rng.uniform(-lim, lim, size=shp), dtype=dtype))

    # ------------------------------------------------------------------------------------------------------------------
    def leaves(self):
        """fresh autograd leaves over the current parameter values (views of the flat buffer are not leaves)"""
        return {k: v.detach().requires_grad_(True) for k, v in self.p.items()}

    def forward(self, x, mask=None, params=None):
        """x [B, X, Y, Z, C] -> [B, 1] (no activation on the output).  mask (mask_input): multiplied onto x (:487)."""
        p = self.p if params is None else params
        if self.mask_input:
            assert mask is not None, 'this discriminator was built with mask_input=True'
            x = x * mask.to(x.dtype)
        t = x.permute(0, 4, 1, 2, 3)
        for name, kind, stride in self.layers:
            if kind == 'conv':
                pads = [same_padding(n, 3, stride) for n in t.shape[2:]]
                t = F.pad(t, (pads[2][0], pads[2][1], pads[1][0], pads[1][1], pads[0][0], pads[0][1]))
                t = F.conv3d(t, p[name + '/kernel'].permute(4, 3, 0, 1, 2), p[name + '/bias'], stride=stride)
                t = F.leaky_relu(t, LEAKY_ALPHA)
            elif name == 'dense_1':
                t = t.permute(0, 2, 3, 4, 1).reshape(t.shape[0], -1)          # Flatten(data_format='channels_last')
                t = F.leaky_relu(t @ p[name + '/kernel'] + p[name + '/bias'], LEAKY_ALPHA)
            else:
                t = t @ p[name + '/kernel'] + p[name + '/bias']
        return t

    __call__ = forward

    def state_dict(self):
        return {k: v.detach().cpu().numpy().copy() for k, v in self.p.items()}

    def load_state_dict(self, sd):
        for k, v in self.p.items():
            v.copy_(torch.as_tensor(np.asarray(sd[k]), dtype=self.dtype).view(v.shape))

    def adam_step(self, lr, lr_decay=0., beta1=.9, beta2=.999, eps=1e-7, grad_scale=1.):
        """keras.optimizers.Adam (Keras 2.3.1) on the flat buffers: the fused kernel on the GPU, the same update in torch
        elsewhere (CPU tests)."""
        lr_eff = lr * (1. / (1. + lr_decay * self.iterations)) if lr_decay > 0 else lr
        t = self.iterations + 1
        lr_t = lr_eff * (math.sqrt(1. - beta2 ** t) / (1. - beta1 ** t))
        if self.params.is_cuda and self.dtype == torch.float32:
            lib.ssr_adam_flat(self.params, self.grads, self.adam_m, self.adam_v, self.n_params, lr_t, beta1, beta2, eps,
                              grad_scale, stream_ptr())
        else:
            g = self.grads * grad_scale
            self.adam_m.mul_(beta1).add_(g, alpha=1. - beta1)
            self.adam_v.mul_(beta2).addcmul_(g, g, value=1. - beta2)
            self.params.sub_(lr_t * self.adam_m / (self.adam_v.sqrt() + eps))
        self.iterations = t


def random_weighted_average(real, fake, weights):
    """RandomWeightedAverage (:606-624): ONE uniform weight per batch element, shape [B, 1, 1, 1, 1]."""
    return weights * real + (1. - weights) * fake


def discriminator_loss(disc, real, fake, weights, gradient_penalty_weight=10., mask=None, params=None):
    """build_discriminator_loss (:580-596) -> (loss, parts).  real / fake [B, X, Y, Z, C]; weights [B, 1, 1, 1, 1].
    The gradient norm is taken over the SPATIAL axes only (axis = 1 .. n_dims, :585), i.e. per batch element and channel."""
    avg = random_weighted_average(real, fake, weights).detach().requires_grad_(True)
    # three calls like the reference (:420-426); ONE call on the stacked batch measured 2.5x slower at 160^3 (1758 against 700 ms
    # per discriminator step: cuDNN's double backward at batch 3)
    d_real, d_fake, d_avg = disc(real, mask, params), disc(fake, mask, params), disc(avg, mask, params)
    grads = torch.autograd.grad(d_avg.sum(), avg, create_graph=True)[0]              # K.gradients(discriminator_av, samples)
    norm = torch.sqrt(torch.sum(grads * grads, dim=tuple(range(1, avg.dim() - 1))))
    penalty = gradient_penalty_weight * (1. - norm) ** 2
    w_real, w_fake, gp = (-d_real).mean(), d_fake.mean(), penalty.mean()
    return w_real + w_fake + gp, (w_real.detach(), w_fake.detach(), gp.detach())


def wasserstein_generator_term(disc, pred, mask=None):
    """w_loss of build_generator_loss (:541): mean(-D(prediction)) and its gradient w.r.t. the prediction."""
    x = pred.detach().clone().requires_grad_(True)
    with torch.enable_grad(), fp32_convs():
        w = (-disc(x, mask)).mean()
        (dx,) = torch.autograd.grad(w, x)
    return w.detach(), dx


class AdversarialUNet3D(SegRegularisedUNet3D):
    """The U-Net while it is fine-tuned against a discriminator.  loss_and_grad computes build_generator_loss (:511-577):
    l1_weight * L1 + discr_weight * mean(-D(pred)) [+ dice_weight * Dice], l1_weight = 1 - discr_weight [- dice_weight].
    `seg` (optional SegRegulariser) must have been built with rel_weight = dice_weight.  Set `seg_labels` (deformed label
    map of the batch) before the step when a segmentation regulariser or a discriminator mask is used."""

    def __init__(self, *args, disc=None, discr_weight=0.01, mask_lut=None, **kwargs):
        super().__init__(*args, **kwargs)
        assert self.nb_labels == 1, 'the discriminator judges one output channel'
        self.disc, self.discr_weight = disc, float(discr_weight)
        self.l1_weight = 1. - self.discr_weight - (self.seg.rel_weight if self.seg is not None else 0.)
        self.mask_lut = mask_lut                      # float32 [max label + 1] on the device, or None
        self.last_terms = None

    def mask_of(self, labels):
        """layers.ConvertLabels(generation_labels, labels_to_mask) on the deformed label map -> [B, X, Y, Z, 1]"""
        if self.mask_lut is None:
            return None
        return self.mask_lut[labels.long()].unsqueeze(-1)

    def _head(self, target, metric, residual, loss_cropping, train):
        UNet3D._head(self, target, metric, residual, loss_cropping, train)
        if not train or self.disc is None:
            return
        name = 'unet_likelihood'
        # image term: everything the head just wrote is linear in the loss weight
        for t in (self.loss_buf, self.dbn_dec[0], self.g[name + '/kernel'], self.g[name + '/bias']):
            t.mul_(self.l1_weight)
        image_term = self.loss_buf.clone()
        if self.seg is not None:
            assert self.seg_labels is not None, 'set seg_labels (deformed label map of this batch) before the step'
            self.seg.add_loss_and_grad(self, self.seg_labels, residual)
        mask = self.mask_of(self.seg_labels) if self.mask_lut is not None else None
        w, dpred = wasserstein_generator_term(self.disc, self.pred.view(self.B, *self.dims, 1), mask)
        self.loss_buf.add_(w.double() * self.discr_weight)
        e = (dpred * self.discr_weight).reshape(-1).contiguous()
        lib.ssr_head_extra_grad(self._feat_src, self._feat_stats, self.p[name + '/kernel'], e, self.nvox[0], self.feats[0],
                                self.dbn_dec[0], self.g[name + '/kernel'], self.g[name + '/bias'], stream_ptr())
        self.last_terms = (image_term, w)


class AdversarialEngine:
    """The two steps the reference alternates (:440-459) on one TrainingEngine (generator + U-Net + Adam + the data-parallel
    exchange) and one Discriminator."""

    def __init__(self, engine, disc, lr_discriminator=1e-4, lr_decay=0., gradient_penalty_weight=10., seed=0):
        self.engine, self.disc = engine, disc
        self.lr_d, self.lr_decay, self.gp_weight = float(lr_discriminator), float(lr_decay), float(gradient_penalty_weight)
        self.gen_w = torch.Generator(device='cpu').manual_seed(int(seed) * 7919 + 13 + engine.rank)

    def _batch(self, labels, means, stds, real_image):
        from .draws import sample_draws
        e = self.engine
        draws = sample_draws(e.rng, e.plan, e.B)
        image, target = e.gen.run(labels, means, stds, draws, real_image=real_image, seed=e.seed)
        return image, target

    def discriminator_step(self, labels, means, stds, real_image=None):
        """one `discriminator_model.train_on_batch` (:449): the U-Net is frozen (batch-statistics forward, no moving-average
        update, no weight update), the discriminator takes one Adam step on w_real + w_fake + gradient penalty."""
        e, d = self.engine, self.disc
        image, target = self._batch(labels, means, stds, real_image)
        fake = e.net.forward_frozen(image).detach().clone()
        real = target.view(fake.shape)
        mask = e.net.mask_of(e.gen.labels) if getattr(e.net, 'mask_lut', None) is not None else None
        weights = torch.rand((e.B, 1, 1, 1, 1), generator=self.gen_w).to(fake.device)
        leaves = d.leaves()
        names = list(leaves)
        with fp32_convs():
            loss, _ = discriminator_loss(d, real, fake, weights, self.gp_weight, mask, leaves)
            grads = torch.autograd.grad(loss, [leaves[k] for k in names])
        d.grads.zero_()
        for k, g in zip(names, grads):
            d.g[k].copy_(g)
        scale = 1.
        if e.world > 1:
            import torch.distributed as dist
            dist.all_reduce(d.grads)
            scale = 1. / e.world
        d.adam_step(self.lr_d, self.lr_decay, grad_scale=scale)
        e.steps += 1                                           # the augmentation counters advance with every batch drawn
        return loss.detach()

    def generator_step(self, labels, means, stds, real_image=None):
        """one `generator_model.train_on_batch` (:456): discriminator frozen, the U-Net takes one Adam step on
        build_generator_loss."""
        return self.engine.train_step(labels, means, stds, real_image=real_image)
