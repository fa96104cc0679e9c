"""Segmentation-regularised loss of the reference (SynthSR/metrics_model.py:136-215 `add_seg_loss_to_model`)
on the B200 engine:  total = image loss + rel_weight * soft Dice(one-hot deformed labels, frozen segmentation U-Net(prediction)).

    FrozenUNet3D            the second U-Net (softmax head): forward with batch-statistics BatchNorm, no moving-average
                            updates, data-gradient-only backward down to its single input channel -- no weight gradients.
    SegRegulariser          input normalisation, softmax + Dice kernels (csrc/seg_loss.cu), the extra gradient through the main
                            network's head.
    SegRegularisedUNet3D    the main network: its head step also runs the regulariser before the backward chain starts.

STATUS: validated on a B200 in round 2 (tests/test_seg_loss_gpu.py: loss 1e-3, gradients 1e-2 against the float64 oracle); nothing on the plain training path
imports this module.  Oracle: oracle/unet.py:seg_regularised_loss (pinned by executing the reference's own function);
GPU tests: tests/test_seg_loss_gpu.py.
"""
import ctypes

import numpy as np
import torch

from ._lib import lib, stream_ptr
from .unet import BN_EPS, UNet3D


class FrozenUNet3D(UNet3D):
    def __init__(self, input_shape, nb_labels, **kw):
        super().__init__(input_shape, nb_labels=nb_labels, **kw)
        self.materialise_feat = True       # the 1x1x1 softmax head reads the normalised feature tensor
        self.head_bn_sums = False
        self.overlap_wgrad = False         # nothing to overlap: there are no weight gradients
        self.logits = torch.empty((self.nvox[0], self.nb_labels), dtype=torch.float32, device=self.device)
        self.dx = torch.empty((self.nvox[0], self.cin), dtype=torch.float32, device=self.device)

    # Keras drops the moving-average updates of a non-trainable layer but still normalises with the batch statistics while
    # fitting (BatchNormalization.call of Keras 2.3.1 does not look at `trainable`): momentum 1 keeps the moving statistics
    def _bn_stats(self, bn, x, nvox, C, stats, training, have_sums=False):
        st = stream_ptr()
        if training and have_sums:
            lib.ssr_bn_finalize(self.sums, nvox, C, self.p[bn + '/gamma'], self.p[bn + '/beta'],
                                self.moving[bn + '/moving_mean'], self.moving[bn + '/moving_variance'], BN_EPS, 1., stats, st)
        elif training:
            lib.ssr_bn_stats(x, nvox, C, self.p[bn + '/gamma'], self.p[bn + '/beta'], self.moving[bn + '/moving_mean'],
                             self.moving[bn + '/moving_variance'], BN_EPS, 1., self.sums, stats, st)
        else:
            super()._bn_stats(bn, x, nvox, C, stats, training, have_sums)

    def _wgrad_async(self, *args, **kwargs):       # frozen: no weight gradients
        return

    def forward_logits(self, image):
        """image [B,X,Y,Z,1] -> logits [B*X*Y*Z, S] (pre-softmax output of `unet_likelihood`, models.py:480-481)."""
        feat = self.forward(image, training=True)
        name = 'unet_likelihood'
        lib.ssr_conv3d_fwd_ref(feat, self.feats[0], None, 0, self.p[name + '/kernel'], self.p[name + '/bias'], self.logits,
                               self.B, *self.dims, self.nb_labels, 1, 0, stream_ptr())
        return self.logits

    def backward_input(self, dlogits):
        """gradient w.r.t. the network input given dL/dlogits: the backward chain of UNet3D.loss_and_grad without its loss
        head and without any weight gradient, plus the data gradient of the first (Cin = 1) convolution."""
        self._alloc_bwd()
        st = stream_ptr()
        B, L, F = self.B, self.L, self.feats
        self.grads.zero_()                         # bias / BN parameter gradients are by-products of the fused kernels; unused
        lib.ssr_conv3d_dgrad_ref(dlogits, self.p['unet_likelihood/kernel'], self.wd_scratch, self.dbn_dec[0], B, *self.dims,
                                 F[0], self.nb_labels, 1, st)
        # ---- decoder, shallow to deep ----------------------------------------------------------------------------------
        for l in range(L - 1):
            d = L - 2 - l
            c0, c1n = 'unet_conv_uparm_%d_0' % (L + d), 'unet_conv_uparm_%d_1' % (L + d)
            bn = 'unet_bn_up_%d' % d
            lib.ssr_bn_bwd(self.dbn_dec[l], self.g1[l], self.stats_dec[l], self.nvox[l], F[l], None, 0, 0, 1,
                           self.ga[l], self.g[bn + '/gamma'], self.g[bn + '/beta'], self.g[c1n + '/bias'], self.sums, st)
            if self._k2n_epi_ok(F[l], F[l]):
                self._conv_dgrad(c1n, self.ga[l], self.gb[l], l, F[l], F[l], elu_h=self.g0[l], dbias=self.g[c0 + '/bias'])
            else:
                self._conv_dgrad(c1n, self.ga[l], self.gb[l], l, F[l], F[l])
                lib.ssr_elu_bwd(self.gb[l], 0, 0, self.g0[l], None, self.nvox[l], F[l], self.gb[l], self.g[c0 + '/bias'], st)
            tgt = self.dbn_dec[l + 1] if l + 1 <= L - 2 else self.dbn_bott
            if l in self.up_levels:
                self._conv_dgrad_up(c0, l, self.gb[l], tgt)
            else:
                self._conv_dgrad(c0, self.gb[l], self.dcat[l], l, F[l] + F[l + 1], F[l])
                lib.ssr_upsample_bwd(self.dcat[l], F[l] + F[l + 1], F[l], B, *self.ldims[l + 1], F[l + 1], tgt, st)
        # ---- encoder, deep to shallow ----------------------------------------------------------------------------------
        for l in range(L - 1, -1, -1):
            c0, c1n, bn = 'unet_conv_downarm_%d_0' % l, 'unet_conv_downarm_%d_1' % l, 'unet_bn_down_%d' % l
            if l == L - 1:
                lib.ssr_bn_bwd(self.dbn_bott, self.h1[l], self.stats_enc[l], self.nvox[l], F[l], None, 0, 0, 1,
                               self.ga_e[l], self.g[bn + '/gamma'], self.g[bn + '/beta'], self.g[c1n + '/bias'], self.sums, st)
            else:
                add, add_stride = (self.dskip[l], F[l]) if l in self.up_levels else (self.dcat[l], F[l] + F[l + 1])
                if self.pool_bn_fusion and F[l] % 4 == 0 and 192 % (F[l] // 4) == 0:
                    lib.ssr_pool_bn_bwd(self.dp[l + 1], self.h1[l], self.stats_enc[l], B, *self.ldims[l], F[l], add,
                                        add_stride, 0, 1, self.ga_e[l], self.g[bn + '/gamma'], self.g[bn + '/beta'],
                                        self.g[c1n + '/bias'], self.sums, st)
                else:
                    lib.ssr_maxpool_bwd(self.dp[l + 1], self.h1[l], self.stats_enc[l], B, *self.ldims[l], F[l], self.ga_e[l], st)
                    lib.ssr_bn_bwd(self.ga_e[l], self.h1[l], self.stats_enc[l], self.nvox[l], F[l], add, add_stride, 0, 1,
                                   self.ga_e[l], self.g[bn + '/gamma'], self.g[bn + '/beta'], self.g[c1n + '/bias'], self.sums, st)
            if self._k2n_epi_ok(F[l], F[l]):
                self._conv_dgrad(c1n, self.ga_e[l], self.gb_e[l], l, F[l], F[l], elu_h=self.h0[l], dbias=self.g[c0 + '/bias'])
            else:
                self._conv_dgrad(c1n, self.ga_e[l], self.gb_e[l], l, F[l], F[l])
                lib.ssr_elu_bwd(self.gb_e[l], 0, 0, self.h0[l], None, self.nvox[l], F[l], self.gb_e[l], self.g[c0 + '/bias'], st)
            if l > 0:
                self._conv_dgrad(c0, self.gb_e[l], self.dp[l], l, F[l - 1], F[l])
            else:                                   # the training path never needs this one: exact-fp32 direct kernel, Cin = 1
                lib.ssr_conv3d_dgrad_ref(self.gb_e[0], self.p[c0 + '/kernel'], self.wd_scratch, self.dx, B, *self.dims,
                                         self.cin, F[0], self.k, st)
        return self.dx


def class_tables(generation_labels, segmentation_label_equivalency, gt_by_value=False):
    """metrics_model.py:185-204 -> (cls_of_seg int32[S], gt_value int32[K]): segmentation channel j belongs to class k when
    equivalency[j] == generation_labels[i] for the k-th such i; the ground truth of that class is `labels == i` -- the loop
    INDEX, as the reference writes it (:188), not the label value.  gt_by_value: `labels == generation_labels[i]`, the way
    fine_tuning_with_adversary.py:551 writes the same loop."""
    eq = np.asarray(segmentation_label_equivalency)
    cls = np.full(len(eq), -1, dtype=np.int32)
    gtv = []
    for i, gl in enumerate(np.asarray(generation_labels)):
        idx = np.where(eq == gl)[0]
        if len(idx) > 0:
            if len(idx) > 3:
                raise Exception("uuummm weird that you're merging so many labels...")
            cls[idx] = len(gtv)
            gtv.append(int(gl) if gt_by_value else i)
    if not gtv:
        raise ValueError('segmentation_label_equivalency matches none of the generation labels')
    return cls, np.asarray(gtv, dtype=np.int32)


class SegRegulariser:
    def __init__(self, dims, batchsize, seg_state_dict, n_seg_labels, generation_labels, segmentation_label_equivalency,
                 rel_weight, loss_cropping=None, m=None, M=None, fs_header=False, nb_features=24, nb_levels=5, conv_size=3,
                 feat_mult=2, nb_conv_per_level=2, conv_impl='tc3', device='cuda', gt_by_value=False):
        if fs_header:
            raise NotImplementedError('fs_header_segnet=True (axis swap + flip around the segmentation network, '
                                      'metrics_model.py:157-162) is not implemented')
        self.dims, self.B = [int(d) for d in dims], int(batchsize)
        self.net = FrozenUNet3D(self.dims + [1], n_seg_labels, nb_features=nb_features, nb_levels=nb_levels,
                                conv_size=conv_size, feat_mult=feat_mult, nb_conv_per_level=nb_conv_per_level,
                                batchsize=batchsize, device=device, conv_impl=conv_impl, seed=0)
        # by name, like the reference's load_weights(by_name=True) (training.py:390): tensors the file does not hold keep
        # their initial values; tensors it holds must have the engine's shape
        for k, v in seg_state_dict.items():
            tgt = self.net.p.get(k, self.net.moving.get(k))
            if tgt is not None and tuple(np.shape(v)) != tuple(tgt.shape):
                raise ValueError('segmentation model file: %s has shape %s, the network expects %s' % (k, np.shape(v), tuple(tgt.shape)))
        self.net.load_state_dict(seg_state_dict, strict=False)
        cls, gtv = class_tables(generation_labels, segmentation_label_equivalency, gt_by_value)
        assert len(cls) == n_seg_labels, 'segmentation_label_equivalency must have one entry per segmentation label'
        dev = self.net.device
        self.S, self.K = int(n_seg_labels), int(len(gtv))
        if self.K == 1:
            # with a single matched class the reference does not stack (metrics_model.py:206-207): DiceLoss then treats Z as
            # the label axis and returns a per-slice Dice, not the whole-volume Dice these kernels compute
            raise NotImplementedError('segmentation regulariser with a single matched class (per-slice Dice in the reference)')
        self.cls, self.gtv = torch.from_numpy(cls).to(dev), torch.from_numpy(gtv).to(dev)
        self.rel_weight = float(rel_weight)
        self.use_clip = m is not None
        self.m, self.M = (float(m), float(M)) if self.use_clip else (0., 1.)
        self._crop = None
        if loss_cropping is not None:
            lc = [int(loss_cropping)] * 3 if isinstance(loss_cropping, (int, np.integer)) else [int(v) for v in loss_cropping]
            cb = [int((self.dims[i] - lc[i]) / 2) for i in range(3)]
            self._crop = ((ctypes.c_int * 3)(*lc), (ctypes.c_int * 3)(*cb))          # HOST arrays, kept alive here
        V = self.net.nvox[0]
        self.x = torch.empty((V, 1), dtype=torch.float32, device=dev)
        self.e = torch.empty(V, dtype=torch.float32, device=dev)
        self.dlogits = torch.empty((V, self.S), dtype=torch.float32, device=dev)
        self.sums = torch.zeros(self.B * self.K * 2, dtype=torch.float64, device=dev)

    def _crop_args(self):
        if self._crop is None:
            return None, None
        return ctypes.cast(self._crop[0], ctypes.c_void_p), ctypes.cast(self._crop[1], ctypes.c_void_p)

    def add_loss_and_grad(self, main, labels, residual=None):
        """main: the U-Net being trained, right after its head step (main.pred, main.loss_buf, main.dbn_dec[0] and the head
        gradients hold the image-loss part); labels: int32 [B, X, Y, Z] deformed label map (`segmentation_target`)."""
        assert main.nb_labels == 1, 'the segmentation network takes one channel (training.py:376)'
        assert list(labels.shape) == [self.B] + self.dims and labels.dtype == torch.int32 and labels.is_contiguous()
        st = stream_ptr()
        V, C = self.net.nvox[0], main.feats[0]
        image, cin, res_c = (main._image, main.cin, int(residual[0])) if residual is not None else (None, 0, 0)
        cs, cb = self._crop_args()
        lib.ssr_seg_input(main.pred, image, cin, res_c, int(self.use_clip), self.m, self.M, self.x, V, st)
        logits = self.net.forward_logits(self.x.view(self.B, *self.dims, 1))
        lib.ssr_softmax_dice_sums(logits, self.S, labels, self.cls, self.gtv, self.K, self.B, *self.dims, cs, cb, self.sums, st)
        lib.ssr_dice_finalize(self.sums, self.B, self.K, self.rel_weight, main.loss_buf, st)
        lib.ssr_softmax_dice_grad(logits, self.S, labels, self.cls, self.gtv, self.K, self.B, *self.dims, cs, cb, self.sums,
                                  self.rel_weight, self.dlogits, st)
        dx = self.net.backward_input(self.dlogits)
        lib.ssr_seg_input_bwd(main.pred, image, cin, res_c, int(self.use_clip), self.m, self.M, dx, self.e, V, st)
        name = 'unet_likelihood'
        lib.ssr_head_extra_grad(main._feat_src, main._feat_stats, main.p[name + '/kernel'], self.e, V, C, main.dbn_dec[0],
                                main.g[name + '/kernel'], main.g[name + '/bias'], st)


class SegRegularisedUNet3D(UNet3D):
    """the network being trained when a segmentation regulariser is attached: `seg_labels` (the generator's deformed label
    map of the current batch) must be set before loss_and_grad."""

    def __init__(self, *args, seg=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.seg = seg
        self.seg_labels = None
        self.head_bn_sums = False          # the head's algebraic BatchNorm reductions do not include the extra gradient

    def _head(self, target, metric, residual, loss_cropping, train):
        super()._head(target, metric, residual, loss_cropping, train)
        if train and self.seg is not None:
            assert self.seg_labels is not None, 'set seg_labels (deformed label map of this batch) before the step'
            self.seg.add_loss_and_grad(self, self.seg_labels, residual)
