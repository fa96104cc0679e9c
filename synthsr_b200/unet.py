"""B200 U-Net training engine: ext/neuron/models.py `unet` (conv_enc :256-360 + conv_dec :363-498) + the loss of
SynthSR/metrics_model.py + Keras Adam, executed as CUDA kernels through the C ABI (include/synthsr_b200.h).

Parameters live in ONE flat float32 buffer (so the data-parallel step is a single all-reduce of one flat gradient
buffer and a single fused Adam launch); views into it carry the Keras layer names and Keras layouts
(kernels (k,k,k,Cin,Cout)), which is what checkpoints store.

conv_impl = 'tc3' : tcgen05 convolutions, forward COMPENSATED to fp32-class accuracy on every layer but the first
                    convolution of the last decoder level -- both operands split into two pieces, the cross terms as extra
                    K-chunks of the same implicit GEMM: x1 w1 + x2 w1 + x1 w2 in bf16 ("bf16x3", the default scheme), a TF32
                    main term + one bf16 correction chain on the 24-channel layers ("hybrid"), or three TF32 chains
                    ("tf32x3", the first implementation) -- backward in plain TF32.  The mode that meets the 1e-3 parity bar
                    on the prediction / loss and 1e-2 on the gradients (tests/test_unet_parity_gpu.py) -- the default.
conv_impl = 'tc'  : plain TF32 everywhere (conv_tc.cu): ~2.9e-4 per convolution, 2-3e-3 on the prediction of a randomly
                    initialised net (scripts/tf32_error_emulation.py) -- throughput mode, outside the parity bar
conv_impl = 'ref' : exact fp32 CUDA-core convolutions (unet_kernels.cu)             -- cross-check
"""
import contextlib
import math
import os
import re
from collections import OrderedDict

import numpy as np
import torch

from ._lib import lib, stream_ptr

BN_EPS = 1e-3
BN_MOMENTUM = 0.99


def layer_specs(cin, nb_features=24, nb_levels=5, feat_mult=2, nb_conv_per_level=2, nb_labels=1):
    """[(keras_name, kind, cin, cout)] in graph order; names as in the reference's .h5 files."""
    specs, c, enc = [], cin, []
    for level in range(nb_levels):
        f = int(np.round(nb_features * feat_mult ** level))
        for j in range(nb_conv_per_level):
            specs.append(('unet_conv_downarm_%d_%d' % (level, j), 'conv', c, f))
            c = f
        specs.append(('unet_bn_down_%d' % level, 'bn', f, f))
        enc.append(f)
    for level in range(nb_levels - 1):
        f = int(np.round(nb_features * feat_mult ** (nb_levels - 2 - level)))
        c = enc[nb_levels - 2 - level] + c
        for j in range(nb_conv_per_level):
            specs.append(('unet_conv_uparm_%d_%d' % (nb_levels + level, j), 'conv', c, f))
            c = f
        specs.append(('unet_bn_up_%d' % level, 'bn', f, f))
    specs.append(('unet_likelihood', 'conv1', c, nb_labels))
    return specs


def keras_layer_order(nb_levels=5, nb_conv_per_level=2, prefix='unet'):
    """every Keras layer name of ext/neuron/models.py `unet` in model.layers order (weightless ones included), the way
    `layer_names` lists them in the reference's .h5 files (models.py:281, 314, 350, 355, 425, 433, 442, 476, 480, 493)."""
    names = ['%s_input' % prefix]
    for level in range(nb_levels):
        names += ['%s_conv_downarm_%d_%d' % (prefix, level, j) for j in range(nb_conv_per_level)]
        names.append('%s_bn_down_%d' % (prefix, level))
        if level < nb_levels - 1:
            names.append('%s_maxpool_%d' % (prefix, level))
    for level in range(nb_levels - 1):
        names += ['%s_up_%d' % (prefix, nb_levels + level), '%s_merge_%d' % (prefix, nb_levels + level)]
        names += ['%s_conv_uparm_%d_%d' % (prefix, nb_levels + level, j) for j in range(nb_conv_per_level)]
        names.append('%s_bn_up_%d' % (prefix, level))
    return names + ['%s_likelihood' % prefix, '%s_prediction' % prefix]


class UNet3D:
    def __init__(self, input_shape, nb_features=24, nb_levels=5, conv_size=3, nb_labels=1, feat_mult=2,
                 nb_conv_per_level=2, batchsize=1, device='cuda', conv_impl='tc3', seed=None):
        assert nb_conv_per_level == 2, 'the reference training path uses nb_conv_per_level=2 (SynthSR/training.py:75)'
        assert conv_size == 3 or conv_impl == 'ref'
        self.B = int(batchsize)
        self.dims = [int(s) for s in input_shape[:3]]
        self.cin = int(input_shape[3])
        self.L = int(nb_levels)
        self.k = int(conv_size)
        self.nb_labels = int(nb_labels)
        assert conv_impl in ('tc', 'tc3', 'ref'), conv_impl
        # compensated forward: {regex over layer names: level}; level 3 = full hi/lo split of activations and weights,
        # level 2 = activations only.  SSR_COMP='regex:level[,regex:level...]' overrides (experiments).
        self.comp = []
        if conv_impl == 'tc3':
            # every convolution but the first one of the last decoder level (72 -> 24 at full resolution: 30 % of the
            # forward FLOPs, and its rounding is neither amplified by later layers nor the last thing before the head)
            last = 'uparm_%d_0' % (2 * int(nb_levels) - 2)
            self.comp = [(re.compile(r'^(?!.*%s).*_conv_' % last), 3)]
            conv_impl = 'tc'
        if os.environ.get('SSR_COMP') and conv_impl == 'tc':
            self.comp = [(re.compile(r.rsplit(':', 1)[0]), int(r.rsplit(':', 1)[1])) for r in os.environ['SSR_COMP'].split(',')]
        # momentum of the moving-statistics update of a training-mode forward; 1.0 = leave them alone (a forward of the network
        # while it is frozen, e.g. under the discriminator steps of the adversarial fine-tuner: Keras drops the updates of a
        # non-trainable layer but still normalises with the batch statistics)
        self.bn_momentum = BN_MOMENTUM
        self._lo = None                 # scratch for the TF32 residual x_lo of the convolution being run (stream ordered)
        self._lo_src = None             # (data_ptr, nvox, channels) of the tensor whose bf16 split _lo currently holds
        self._ksplit_cache = {}
        # 'bf16x3' (default): x = x1 + x2, w = w1 + w2 in 8-bit pieces, x1 w1 + x2 w1 + x1 w2 as three bf16 K-chunks per 64
        #     input channels = 1.5 TF32 chains (generic / parity kernels from 48 channels on; the k2n layers and 24-channel
        #     operands run 'hybrid');
        # 'hybrid': x_hi w_hi in TF32 + the two correction terms as one bf16 MMA chain (2 chains per convolution);
        # 'tf32x3': all three terms in TF32 (3 chains) -- the first implementation, kept as a cross-check
        self.comp_scheme = os.environ.get('SSR_COMP_SCHEME', 'bf16x3')
        assert self.comp_scheme in ('bf16x3', 'hybrid', 'tf32x3'), self.comp_scheme
        self.conv_impl = conv_impl
        self.wgrad_tc = conv_impl == 'tc'
        self.prof = None          # list of (kind, flops, start_event, end_event) when profiling is enabled
        self.overlap_wgrad = os.environ.get('SSR_NO_WGRAD_OVERLAP') is None
        self.fwd_k2n = os.environ.get('SSR_NO_FWD_K2N') is None
        self.fwd_k2n_parts = self.fwd_k2n and os.environ.get('SSR_NO_FWD_K2N_PARTS') is None
        self.materialise_feat = os.environ.get('SSR_MATERIALISE_FEAT') is not None
        # BN statistics / ELU backward of the full-resolution 24-channel layers inside the k2n convolution epilogues
        self.pool_bn_fusion = os.environ.get('SSR_NO_POOL_BN_FUSION') is None    # MaxPool + BN backward in two passes
        self.head_bn_sums = os.environ.get('SSR_NO_HEAD_BN_SUMS') is None        # BN-backward reductions from the head
        # decoder convolutions over the upsampled tensor as 8 parity classes of effective 2x2x2 kernels on the
        # low-resolution tensor (conv3d_tc_up_kernel): levels whose low-resolution grid is at least up_min_dim wide
        self.up_parity = conv_impl == 'tc' and os.environ.get('SSR_NO_UP_PARITY') is None
        self.up_min_dim = int(os.environ.get('SSR_UP_MIN_DIM', '16'))
        # last decoder level: forward of the upsampled part in the k2n layout (conv3d_tc_up_k2n_kernel)
        self.up_k2n = os.environ.get('SSR_NO_UP_K2N') is None
        # weight gradient of those layers from the low-resolution tensor too (the upsampled tensor is never materialised)
        self.up_wgrad = self.up_parity and os.environ.get('SSR_NO_UP_WGRAD') is None
        self.epi_fusion = self.fwd_k2n and os.environ.get('SSR_NO_EPI_FUSION') is None
        self.epi_fusion_generic = os.environ.get('SSR_NO_EPI_FUSION_GENERIC') is None   # same for the levels below
        self._side, self._side_busy, self._hp, self._pack_event = None, False, None, None
        self.device = torch.device(device)
        for d in self.dims:
            if d % (2 ** (self.L - 1)) != 0:
                raise ValueError('spatial dims %s must be divisible by %d (UpSampling/concatenate would not match; the '
                                 'reference enforces output_div_by_n=2**n_levels)' % (self.dims, 2 ** (self.L - 1)))
        self.specs = layer_specs(self.cin, nb_features, nb_levels, feat_mult, nb_conv_per_level, nb_labels)
        self.feats = [int(np.round(nb_features * feat_mult ** l)) for l in range(self.L)]
        # ---- flat parameter / gradient / Adam buffers with named views -------------------------------------------
        # Offsets follow the order in which the BACKWARD pass completes the gradients (head first, first encoder level
        # last), so that at any point of the backward the finished gradients are a contiguous PREFIX of the buffer: the
        # data-parallel exchange all-reduces that prefix while the shallow levels are still being differentiated
        # (trainer.GradientExchange).  Names / shapes keep the graph order (what checkpoints list).
        offs, off = {}, 0
        self.level_end = {}              # encoder level l -> end offset of everything complete once level l is done
        # (the 1x1x1 head keeps the END of the buffer: its odd size would break the 16-byte alignment of what follows)
        for name, kind, ci, co in list(reversed(self.specs[:-1])) + [self.specs[-1]]:
            if kind in ('conv', 'conv1'):
                k = self.k if kind == 'conv' else 1
                offs[name + '/kernel'] = off; off += k ** 3 * ci * co
                offs[name + '/bias'] = off; off += co
                if name.startswith('unet_conv_downarm_') and name.endswith('_0'):
                    self.level_end[int(name.split('_')[3])] = off
            else:
                offs[name + '/gamma'] = off; off += co
                offs[name + '/beta'] = off; off += co
        self.layout = OrderedDict()
        for name, kind, ci, co in self.specs:
            if kind in ('conv', 'conv1'):
                k = self.k if kind == 'conv' else 1
                self.layout[name + '/kernel'] = (offs[name + '/kernel'], (k, k, k, ci, co))
                self.layout[name + '/bias'] = (offs[name + '/bias'], (co,))
            else:
                self.layout[name + '/gamma'] = (offs[name + '/gamma'], (co,))
                self.layout[name + '/beta'] = (offs[name + '/beta'], (co,))
        self.n_params = off
        dev = self.device
        n_mv = sum(2 * co for name, kind, ci, co in self.specs if kind == 'bn')
        self.params = torch.zeros(off, dtype=torch.float32, device=dev)
        # exchange buffer of the data-parallel step: [gradients | BN moving statistics | loss], all-reduced IN PLACE
        mv0 = (off + 3) // 4 * 4                         # 16-byte aligned start of the moving statistics
        self.comm = torch.zeros(mv0 + n_mv + 1, dtype=torch.float32, device=dev)
        self.grads = self.comm[:off]
        self.adam_m = torch.zeros(off, dtype=torch.float32, device=dev)
        self.adam_v = torch.zeros(off, dtype=torch.float32, device=dev)
        self.iterations = 0
        self.grads_ready_hook = None     # callable(level): the gradients of encoder level `level` and of everything
                                         # differentiated before it (a prefix of `grads`) have been enqueued
        self.moving = OrderedDict()
        o = mv0
        for name, kind, ci, co in self.specs:
            if kind == 'bn':
                self.moving[name + '/moving_mean'] = self.comm[o:o + co]; o += co
                self.moving[name + '/moving_variance'] = self.comm[o:o + co]; o += co
                self.moving[name + '/moving_variance'].fill_(1.)
        self.p = {n: self.params[o:o + int(np.prod(s))].view(s) for n, (o, s) in self.layout.items()}
        self.g = {n: self.grads[o:o + int(np.prod(s))].view(s) for n, (o, s) in self.layout.items()}
        self.init_weights(seed)
        self._alloc()

    # -------------------------------------------------------------------------------------------------------------
    def init_weights(self, seed=None):
        """glorot_uniform kernels, zero biases, gamma=1, beta=0 (Keras defaults used by models.py:297-351)."""
        rng = np.random.default_rng(seed)
        host = np.zeros(self.n_params, dtype=np.float32)
        for name, kind, ci, co in self.specs:
            if kind in ('conv', 'conv1'):
                k = self.k if kind == 'conv' else 1
                limit = math.sqrt(6.0 / (k ** 3 * ci + k ** 3 * co))
                o, s = self.layout[name + '/kernel']
                host[o:o + int(np.prod(s))] = rng.uniform(-limit, limit, size=int(np.prod(s))).astype(np.float32)
            else:
                o, s = self.layout[name + '/gamma']
                host[o:o + co] = 1.
        self.params.copy_(torch.from_numpy(host))
        for n, t in self.moving.items():
            t.fill_(0. if n.endswith('mean') else 1.)
        self.adam_m.zero_(); self.adam_v.zero_(); self.iterations = 0
        self._packed_dirty = True

    def state_dict(self):
        sd = OrderedDict((n, self.p[n].detach().cpu().numpy().copy()) for n in self.layout)
        for n, t in self.moving.items():
            sd[n] = t.detach().cpu().numpy().copy()
        return sd

    def load_state_dict(self, sd, strict=True):
        for n in self.layout:
            if n in sd:
                self.p[n].copy_(torch.as_tensor(np.asarray(sd[n], dtype=np.float32)).view(self.p[n].shape))
            elif strict:
                raise KeyError(n)
        for n in self.moving:
            if n in sd:
                self.moving[n].copy_(torch.as_tensor(np.asarray(sd[n], dtype=np.float32)))
            elif strict:
                raise KeyError(n)
        self._packed_dirty = True

    # -------------------------------------------------------------------------------------------------------------
    def _alloc(self):
        B, L, F, dev = self.B, self.L, self.feats, self.device
        f32 = torch.float32
        self.ldims = [[d // (2 ** l) for d in self.dims] for l in range(L)]
        self.nvox = [B * int(np.prod(d)) for d in self.ldims]

        def buf(l, c):
            return torch.empty((self.nvox[l], c), dtype=f32, device=dev)

        self.inp = [None] + [buf(l, F[l - 1]) for l in range(1, L)]           # pooled BN output feeding level l
        self.h0 = [buf(l, F[l]) for l in range(L)]
        self.h1 = [buf(l, F[l]) for l in range(L)]
        self.u = [None] * (L - 1)                                             # upsampled BN output (decoder input), lazily
        self.g0 = [buf(l, F[l]) for l in range(L - 1)]
        self.g1 = [buf(l, F[l]) for l in range(L - 1)]
        self.up_levels = [l for l in range(L - 1) if self.up_parity and F[l] % 8 == 0 and F[l + 1] % 8 == 0 and
                          F[l + 1] <= 256 and min(self.ldims[l + 1]) >= self.up_min_dim]
        self.vlow = {l: buf(l + 1, F[l + 1]) for l in self.up_levels}            # BN output feeding decoder level l
        self._up = {}
        self._up_valid = False
        self.feat = buf(0, F[0])
        self.pred = torch.empty((self.nvox[0], self.nb_labels), dtype=f32, device=dev)
        self.stats_enc = [torch.empty(4 * F[l], dtype=f32, device=dev) for l in range(L)]
        self.stats_dec = [torch.empty(4 * F[l], dtype=f32, device=dev) for l in range(L - 1)]
        self.sums = torch.zeros(2 * max(F), dtype=torch.float64, device=dev)
        self.loss_buf = torch.zeros(1, dtype=torch.float64, device=dev)
        self._bwd_alloc = False
        self._packed = {}
        self._packed_args = {}
        self._packed_src = {}
        self._packed_dirty = True

    def _alloc_bwd(self):
        if self._bwd_alloc:
            return
        L, F, dev = self.L, self.feats, self.device
        f32 = torch.float32

        def buf(l, c):
            return torch.empty((self.nvox[l], c), dtype=f32, device=dev)

        self.ga = [buf(l, F[l]) for l in range(L)]
        self.gb = [buf(l, F[l]) for l in range(L)]
        # the encoder phase has its own buffers: weight-gradient kernels of the decoder phase may still be reading
        # ga/gb on the side stream when the encoder phase starts (see _wgrad_async)
        self.ga_e = [buf(l, F[l]) for l in range(L)]
        self.gb_e = [buf(l, F[l]) for l in range(L)]
        self.dcat = [None if l in self.up_levels else buf(l, F[l] + F[l + 1]) for l in range(L - 1)]
        self.dskip = {l: buf(l, F[l]) for l in self.up_levels}                # gradient w.r.t. the skip input of level l
        self.dbn_dec = [buf(l, F[l]) for l in range(L - 1)]                   # grad wrt BN output of decoder level l
        self.dbn_bott = buf(L - 1, F[L - 1])                                  # grad wrt BN output of the bottleneck
        self.dp = [None] + [buf(l, F[l - 1]) for l in range(1, L)]            # grad wrt pooled input of level l
        self.gout = torch.empty((self.nvox[0], self.nb_labels), dtype=f32, device=dev)
        maxw = max(int(np.prod(s)) for n, (o, s) in self.layout.items() if n.endswith('kernel'))
        self.wd_scratch = torch.empty(maxw, dtype=f32, device=dev)
        self._bwd_alloc = True

    # -------------------------------------------------------------------------------------------------------------
    # convolution dispatch
    # -------------------------------------------------------------------------------------------------------------
    def _timed(self, kind, l, cin, cout, fn, name=None):
        if self.prof is None:
            return fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        # 5th entry: MMA chains executed per algorithmic one (compensated forward: 1.5 bf16x3, 2 hybrid, 3 in 3xTF32)
        mult = 1
        if name is not None and kind == 'fwd_tc' and self._comp_level(name) == 3:
            mult = {'hybrid': 2, 'bf16x3': 2 if (self._k2n_ok(cin, cout) or cin < 48) else 1.5}.get(self.comp_scheme, 3)
        self.prof.append((kind, 2. * self.k ** 3 * cin * cout * self.nvox[l], e0, e1, mult))

    def _k2n_ok(self, cin, cout):
        return self.fwd_k2n and cin % 8 == 0 and cin <= 32 and cout in (24, 32)

    def _ksplit(self, l, cin, cout, comp):
        """split-K factor the generic kernel picks for a cin -> cout convolution at level l without fused epilogues"""
        key = (l, cin, cout, comp)
        if key not in self._ksplit_cache:
            self._ksplit_cache[key] = lib.ssr_conv3d_fwd_tc_ksplit(cin, 0, cout, self.B, *self.ldims[l], comp)
        return self._ksplit_cache[key]

    def _k2n_epi_ok(self, cin, cout, l=None, name=None):
        """BN sums / ELU backward + bias gradient inside the convolution epilogue: k2n kernel for the 24 / 32-channel
        layers, generic kernel (conv3d_tc_kernel<EPI>) for the others -- except on the small deep levels, where the plain
        entry point runs split-K (much faster there) and the BatchNorm sums / ELU' run as their own small passes.
        l: level of the layer; name: the forward layer (its compensation decides the K length), None for a data gradient."""
        if not (self.conv_impl == 'tc' and self.epi_fusion and cin % 8 == 0 and cout % 8 == 0):
            return False
        if self._k2n_ok(cin, cout):
            return True
        if not (self.epi_fusion_generic and not (cin <= 32 and cout <= 32)):
            return False
        if l is not None:
            comp = 0
            if name is not None and self._comp_level(name):
                comp = {'hybrid': 4, 'bf16x3': self._x3(cin)[2]}.get(self.comp_scheme, 3) if self._comp_level(name) == 3 else 3
            if self._ksplit(l, cin, cout, comp) > 1:
                return False
        return True

    def _conv_fwd(self, name, x1, c1, x2, c2, y, l, cout, act=1, stats_sums=None):
        """stats_sums (2*cout doubles): the convolution also accumulates the BatchNorm sums of its output in its epilogue
        (only offered by the caller when _k2n_epi_ok)."""
        tc = self.conv_impl == 'tc' and c1 % 8 == 0 and c2 % 8 == 0
        self._timed('fwd_tc' if tc else 'fwd_ref', l, c1 + c2, cout,
                    lambda: self._conv_fwd_impl(tc, name, x1, c1, x2, c2, y, l, cout, act, stats_sums), name=name)

    def _comp_level(self, name):
        for rx, level in self.comp:
            if rx.search(name):
                return level
        return 0

    def _residual(self, x, n):
        """x_lo = x - rne_tf32(x) of the first n floats of x, in the shared scratch (consumed by the next launch on this
        stream, so one buffer serves every layer)."""
        self._lo_src = None
        lib.ssr_tf32_residual(x, self._lo_buf(n), n, stream_ptr())
        return self._lo

    def _lo_buf(self, n):
        if self._lo is None or self._lo.numel() < n:
            self._lo = torch.empty(max(n, self.nvox[0] * self.feats[0]), dtype=torch.float32, device=self.device)
        return self._lo

    def _split16(self, x, nvox, c, x3=False):
        """x2 = [bf16(x_lo) | bf16(x_hi)] (2c bf16 channels per voxel = the bytes of c floats) in the shared scratch;
        skipped when the kernel that produced x has just written it from its own epilogue (_lo_src).
        x3: the bf16x3 scheme's operand [bf16(x) | bf16(x - bf16(x))] instead."""
        key = (x.data_ptr(), nvox, c)
        if self._lo_src == key and not x3:
            self._lo_src = None
            return self._lo
        self._lo_src = None
        (lib.ssr_bf16x3_split if x3 else lib.ssr_tf32_split_bf16)(x, self._lo_buf(nvox * c), nvox, c, stream_ptr())
        return self._lo

    def _fuse_split_ok(self, next_name, c1, cout):
        """the layer `next_name` (c1 -> cout) will run the hybrid compensated forward through the k2n kernel on the tensor
        being produced now: the producer may emit its bf16 split directly (the two full-resolution 24-channel tensors)"""
        return (self.conv_impl == 'tc' and self.comp_scheme in ('hybrid', 'bf16x3') and self._comp_level(next_name) == 3 and
                self._k2n_ok(c1, cout) and os.environ.get('SSR_NO_SPLIT_FUSION') is None)

    def _x3(self, c):
        """(bf16x3?, weight pack mode, compensation level) of a c-channel operand in the generic / parity kernels.  bf16x3
        wins from 48 channels on (3 bf16 chunks against 2 TF32 + 2 bf16); a 24-channel operand fills a 64-channel bf16 chunk
        to 3/8 and is cheaper in the hybrid scheme (24 -> 48 @ 80^3: 0.170 ms hybrid, 0.205 ms bf16x3)."""
        x3 = self.comp_scheme == 'bf16x3' and c >= 48
        return (True, 9, 5) if x3 else (False, 7, 4)

    def _conv_fwd_hybrid(self, name, x1, c1, x2, c2, y, l, cout, act, stats_sums):
        """compensated forward, hybrid scheme: TF32 main term + one bf16 chain for x_lo w_hi + x_hi w_lo"""
        st, B, d = stream_ptr(), self.B, self.ldims[l]
        nv = self.nvox[l]
        bias = self.p[name + '/bias']
        if c2 == 0 and self._k2n_ok(c1, cout):
            whi = self._packed_w(name, 2, c1, 0, cout)
            w16 = self._packed_w(name, 8, c1, 0, cout)
            x16 = self._split16(x1, nv, c1)
            lib.ssr_conv3d_fwd_tc_k2n_part(x1, c1, 0, c1, whi, bias, y, B, *d, cout, act, 0, 0, st)
            lib.ssr_conv3d_fwd_tc_k2n_bf16(x16, 2 * c1, w16, bias, y, stats_sums, B, *d, cout, act, st)
            return
        # generic kernel: TF32 main term + bf16 correction chain (hybrid, level 4), or three bf16 terms (bf16x3, level 5)
        if c2 == 0:
            x3, pm, lv = self._x3(c1)
            wp = self._packed_w(name, pm, c1, c1, cout)
            lib.ssr_conv3d_fwd_tc_comp(x1, self._split16(x1, nv, c1, x3), c1, wp, bias, y, stats_sums, B, *d, cout, act, 0, lv, st)
            return
        assert stats_sums is None
        x3, pm, lv = self._x3(c1)
        wp1 = self._packed_w(name, pm, c1 + c2, c1, cout, tag='c0')
        lib.ssr_conv3d_fwd_tc_comp(x1, self._split16(x1, nv, c1, x3), c1, wp1, None, y, None, B, *d, cout, 0, 0, lv, st)
        x3, pm, lv = self._x3(c2)
        wp2 = self._packed_w(name, pm, c1 + c2, (c1 << 12) | c2, cout, tag='c1')
        lib.ssr_conv3d_fwd_tc_comp(x2, self._split16(x2, nv, c2, x3), c2, wp2, bias, y, None, B, *d, cout, act, 1, lv, st)

    def _conv_fwd_comp(self, level, name, x1, c1, x2, c2, y, l, cout, act, stats_sums):
        """compensated forward of one layer (see the module docstring); the concatenated input of a decoder level that
        does not take the parity path is the sum of its two parts."""
        if level == 3 and self.comp_scheme in ('hybrid', 'bf16x3'):
            return self._conv_fwd_hybrid(name, x1, c1, x2, c2, y, l, cout, act, stats_sums)
        st, B, d = stream_ptr(), self.B, self.ldims[l]
        nv = self.nvox[l]
        bias = self.p[name + '/bias']
        if c2 == 0 and self._k2n_ok(c1, cout):
            # full-resolution 24-channel layers: three passes of the k2n kernel (its 27 taps stay resident per pass)
            whi = self._packed_w(name, 2, c1, 0, cout)
            lo = self._residual(x1, nv * c1)
            lib.ssr_conv3d_fwd_tc_k2n_part(x1, c1, 0, c1, whi, bias, y, B, *d, cout, act, 0, 0, st)
            last_hi = level == 2
            if last_hi and stats_sums is not None:
                lib.ssr_conv3d_fwd_tc_k2n_part_stats(lo, c1, 0, c1, whi, bias, y, stats_sums, B, *d, cout, act, 1, st)
            else:
                lib.ssr_conv3d_fwd_tc_k2n_part(lo, c1, 0, c1, whi, bias, y, B, *d, cout, act, 1, 1 if last_hi else 0, st)
            if level == 3:
                wlo = self._packed_w(name, 6, c1, 0, cout)
                if stats_sums is not None:
                    lib.ssr_conv3d_fwd_tc_k2n_part_stats(x1, c1, 0, c1, wlo, bias, y, stats_sums, B, *d, cout, act, 1, st)
                else:
                    lib.ssr_conv3d_fwd_tc_k2n_part(x1, c1, 0, c1, wlo, bias, y, B, *d, cout, act, 1, 1, st)
            return
        if c2 == 0:
            wp = self._packed_w(name, 5, c1, c1, cout)
            lib.ssr_conv3d_fwd_tc_comp(x1, self._residual(x1, nv * c1), c1, wp, bias, y, stats_sums, B, *d, cout, act, 0,
                                       level, st)
            return
        assert stats_sums is None
        wp1 = self._packed_w(name, 5, c1 + c2, c1, cout, tag='c0')
        lib.ssr_conv3d_fwd_tc_comp(x1, self._residual(x1, nv * c1), c1, wp1, None, y, None, B, *d, cout, 0, 0, level, st)
        wp2 = self._packed_w(name, 5, c1 + c2, (c1 << 12) | c2, cout, tag='c1')
        lib.ssr_conv3d_fwd_tc_comp(x2, self._residual(x2, nv * c2), c2, wp2, bias, y, None, B, *d, cout, act, 1, level, st)

    def _conv_fwd_impl(self, tc, name, x1, c1, x2, c2, y, l, cout, act, stats_sums=None):
        st = stream_ptr()
        d = self.ldims[l]
        level = self._comp_level(name) if tc else 0
        if level:
            return self._conv_fwd_comp(level, name, x1, c1, x2, c2, y, l, cout, act, stats_sums)
        if stats_sums is not None and self._k2n_ok(c1, cout):
            assert tc and c2 == 0
            lib.ssr_conv3d_fwd_tc_k2n_stats(x1, c1, self._packed_w(name, 2, c1, 0, cout), self.p[name + '/bias'], y,
                                            stats_sums, self.B, *d, cout, act, st)
        elif stats_sums is not None:
            assert tc
            lib.ssr_conv3d_fwd_tc_stats(x1, c1, x2, c2, self._packed_w(name, 0, c1, c2, cout), self.p[name + '/bias'], y,
                                        stats_sums, self.B, *d, cout, act, st)
        elif tc and self.fwd_k2n and c2 == 0 and c1 <= 32 and cout <= 32:
            # full-resolution 24-channel layers: d2 taps in the MMA N dimension (conv3d_tc_k2n_kernel)
            lib.ssr_conv3d_fwd_tc_k2n(x1, c1, self._packed_w(name, 2, c1, 0, cout), self.p[name + '/bias'], y,
                                      self.B, *d, cout, act, st)
        elif tc and self.fwd_k2n_parts and c2 > 0 and c1 <= 32 and cout <= 32 and c2 % 8 == 0:
            # concatenated input of the last decoder level: sum over <= 32-channel parts, each through the k2n kernel
            parts = [(x1, c1, 0, c1, 0)] + [(x2, c2, o, min(32, c2 - o), c1 + o) for o in range(0, c2, 32)]
            for i, (src, ctot, c0, cn, coff) in enumerate(parts):
                wp = self._packed_w(name, 4, c1 + c2, (coff << 8) | cn, cout, tag=i)
                lib.ssr_conv3d_fwd_tc_k2n_part(src, ctot, c0, cn, wp, self.p[name + '/bias'], y, self.B, *d, cout, act,
                                               1 if i > 0 else 0, 1 if i == len(parts) - 1 else 0, st)
        elif tc:
            lib.ssr_conv3d_fwd_tc(x1, c1, x2, c2, self._packed_w(name, 0, c1, c2, cout), self.p[name + '/bias'], y,
                                  self.B, *d, cout, act, st)
        else:
            lib.ssr_conv3d_fwd_ref(x1, c1, x2, c2, self.p[name + '/kernel'], self.p[name + '/bias'], y, self.B, *d,
                                   cout, self.k, act, st)

    def _conv_fwd_up(self, name, l, act=1):
        """decoder convolution 0 of level l on [skip h1[l], upsample(vlow[l])] without the upsampled tensor: the parity
        kernel writes the partial sums of the upsampled part, the skip convolution accumulates + bias + ELU."""
        F = self.feats
        self._timed('fwd_tc', l, F[l] + F[l + 1], F[l], lambda: self._conv_fwd_up_impl(name, l, act), name=name)

    def _conv_fwd_up_impl(self, name, l, act):
        st, F, B = stream_ptr(), self.feats, self.B
        u = self._up_state(l)
        level = self._comp_level(name)
        if level == 3 and self.comp_scheme in ('hybrid', 'bf16x3'):
            x3, pm, lv = self._x3(F[l + 1])
            lib.ssr_conv3d_fwd_tc_up_comp(self.vlow[l], self._split16(self.vlow[l], self.nvox[l + 1], F[l + 1], x3), F[l + 1],
                                          self._up_packs(l, 'fwd8b' if x3 else 'fwd8h'), self.g0[l], B, *self.ldims[l + 1],
                                          F[l], lv, st)
            x3, pm, lv = self._x3(F[l])
            wp = self._packed_w(name, pm, F[l], F[l], F[l], tag='skipb' if x3 else 'skiph', src=u['wskip'])
            lib.ssr_conv3d_fwd_tc_comp(self.h1[l], self._split16(self.h1[l], self.nvox[l], F[l], x3), F[l], wp,
                                       self.p[name + '/bias'], self.g0[l], None, B, *self.ldims[l], F[l], act, 1, lv, st)
            return
        if level:
            # compensated: parity kernel on [vlow | vlow_lo | vlow], then the skip part accumulates (+ bias + ELU)
            nlow, nsk = self.nvox[l + 1] * F[l + 1], self.nvox[l] * F[l]
            lib.ssr_conv3d_fwd_tc_up_comp(self.vlow[l], self._residual(self.vlow[l], nlow), F[l + 1],
                                          self._up_packs(l, 'fwd8c'), self.g0[l], B, *self.ldims[l + 1], F[l], level, st)
            wp = self._packed_w(name, 5, F[l], F[l], F[l], tag='skipc', src=u['wskip'])
            lib.ssr_conv3d_fwd_tc_comp(self.h1[l], self._residual(self.h1[l], nsk), F[l], wp, self.p[name + '/bias'],
                                       self.g0[l], None, B, *self.ldims[l], F[l], act, 1, level, st)
            return
        if self._up_k2n_ok(l):
            if not self._up_valid:
                self._up_weights_all(st)
            lib.ssr_conv3d_fwd_tc_up_k2n(self.vlow[l], F[l + 1], u['wpk'], self.g0[l], B, *self.ldims[l + 1], F[l], st)
        else:
            lib.ssr_conv3d_fwd_tc_up(self.vlow[l], F[l + 1], self._up_packs(l, 'fwd8'), self.g0[l], B, *self.ldims[l + 1],
                                     F[l], st)
        if self.fwd_k2n and F[l] <= 32:
            wp = self._packed_w(name, 2, F[l], 0, F[l], tag='skip', src=u['wskip'])
            if self._fuse_split_ok(name[:-1] + '1', F[l], F[l]):      # the next convolution's bf16 operand from this epilogue
                lib.ssr_conv3d_fwd_tc_k2n_part_split(self.h1[l], F[l], 0, F[l], wp, self.p[name + '/bias'], self.g0[l],
                                                     self._lo_buf(self.nvox[l] * F[l]), B, *self.ldims[l], F[l], act, 1, st)
                self._lo_src = (self.g0[l].data_ptr(), self.nvox[l], F[l])
                return
            lib.ssr_conv3d_fwd_tc_k2n_part(self.h1[l], F[l], 0, F[l], wp, self.p[name + '/bias'], self.g0[l], B,
                                           *self.ldims[l], F[l], act, 1, 1, st)
        else:
            wp = self._packed_w(name, 0, F[l], 0, F[l], tag='skip', src=u['wskip'])
            lib.ssr_conv3d_fwd_tc_acc(self.h1[l], F[l], None, 0, wp, self.p[name + '/bias'], self.g0[l], B, *self.ldims[l],
                                      F[l], act, st)

    def _conv_dgrad_up(self, name, l, dy, dlow):
        """data gradient of decoder convolution 0 of level l: dskip[l] (w.r.t. the skip input, full resolution) and dlow
        (w.r.t. the low-resolution BN output; UpSampling3D backward included)."""
        F = self.feats
        self._timed('dgrad_tc', l, F[l] + F[l + 1], F[l], lambda: self._conv_dgrad_up_impl(name, l, dy, dlow))

    def _conv_dgrad_up_impl(self, name, l, dy, dlow):
        st, F, B = stream_ptr(), self.feats, self.B
        u = self._up_state(l)
        if self.fwd_k2n and F[l] <= 32:
            wp = self._packed_w(name, 3, F[l], 0, F[l], tag='skip', src=u['wskip'])
            lib.ssr_conv3d_fwd_tc_k2n(dy, F[l], wp, None, self.dskip[l], B, *self.ldims[l], F[l], 0, st)
        else:
            wp = self._packed_w(name, 1, F[l], 0, F[l], tag='skip', src=u['wskip'])
            lib.ssr_conv3d_fwd_tc(dy, F[l], None, 0, wp, None, self.dskip[l], B, *self.ldims[l], F[l], 0, st)
        lib.ssr_conv3d_dgrad_tc_up(dy, F[l], self._up_packs(l, 'dgr8'), dlow, B, *self.ldims[l + 1], F[l + 1], st)

    def _conv_dgrad(self, name, dy, dx, l, cin, cout, elu_h=None, dbias=None):
        """elu_h / dbias: fuse the ELU backward of the layer below (dx *= elu'(elu_h), dbias += column sums of dx)."""
        tc = self.conv_impl == 'tc' and cout % 8 == 0
        self._timed('dgrad_tc' if tc else 'dgrad_ref', l, cin, cout,
                    lambda: self._conv_dgrad_impl(tc, name, dy, dx, l, cin, cout, elu_h, dbias))

    def _conv_dgrad_impl(self, tc, name, dy, dx, l, cin, cout, elu_h=None, dbias=None):
        st = stream_ptr()
        d = self.ldims[l]
        if elu_h is not None and self._k2n_ok(cout, cin):
            assert tc
            lib.ssr_conv3d_dgrad_tc_k2n_elu(dy, cout, self._packed_w(name, 3, cin, 0, cout), elu_h, dx, dbias, self.B,
                                            *d, cin, st)
        elif elu_h is not None:
            assert tc
            lib.ssr_conv3d_dgrad_tc_elu(dy, cout, self._packed_w(name, 1, cin, 0, cout), elu_h, dx, dbias, self.B, *d, cin,
                                        st)
        elif tc and self.fwd_k2n and cin <= 32 and cout <= 32:
            lib.ssr_conv3d_fwd_tc_k2n(dy, cout, self._packed_w(name, 3, cin, 0, cout), None, dx, self.B, *d, cin, 0, st)
        elif tc:
            # data gradient = forward convolution of dy with the flipped / transposed kernel
            lib.ssr_conv3d_fwd_tc(dy, cout, None, 0, self._packed_w(name, 1, cin, 0, cout), None, dx, self.B, *d, cin,
                                  0, st)
        else:
            lib.ssr_conv3d_dgrad_ref(dy, self.p[name + '/kernel'], self.wd_scratch, dx, self.B, *d, cin, cout, self.k,
                                     st)

    def _conv_wgrad(self, name, x1, c1, x2, c2, dy, l, cout):
        tc = self.conv_impl == 'tc' and self._wgrad_tc_ok(c1, c2, cout)
        self._timed('wgrad_tc' if tc else 'wgrad_ref', l, c1 + c2, cout,
                    lambda: self._conv_wgrad_impl(tc, name, x1, c1, x2, c2, dy, l, cout))

    def _conv_wgrad_impl(self, tc, name, x1, c1, x2, c2, dy, l, cout):
        st = stream_ptr()
        d = self.ldims[l]
        if tc:
            nbytes = lib.ssr_conv3d_wgrad_scratch_bytes(c1, c2, cout, self.B, *d)
            if getattr(self, '_wg_scratch', None) is None or self._wg_scratch.numel() * 4 < nbytes:
                self._wg_scratch = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=self.device)
            lib.ssr_conv3d_wgrad_tc(x1, c1, x2, c2, dy, self.g[name + '/kernel'], None,
                                    self._wg_scratch, self._wg_scratch.numel() * 4, self.B, *d, cout, st)
        else:
            lib.ssr_conv3d_wgrad_ref(x1, c1, x2, c2, dy, self.g[name + '/kernel'], None, self.B, *d,
                                     cout, self.k, st)

    def _conv_wgrad_up(self, name, l, dy):
        """weight gradient of decoder convolution 0 of level l: skip channels from h1[l], upsampled channels from the
        low-resolution tensor vlow[l] (gradients of the 8 effective kernels, combined into the 3x3x3 gradient)."""
        F = self.feats
        self._timed('wgrad_tc', l, F[l] + F[l + 1], F[l], lambda: self._conv_wgrad_up_impl(name, l, dy))

    def _conv_wgrad_up_impl(self, name, l, dy):
        st, F, B = stream_ptr(), self.feats, self.B
        u = self._up_state(l)
        if 'gscratch' not in u:
            u['gscratch'] = torch.empty(8 * 27 * F[l + 1] * F[l], dtype=torch.float32, device=self.device)
        dw = self.g[name + '/kernel']
        lib.ssr_conv3d_wgrad_tc_part(self.h1[l], F[l], dy, dw, F[l] + F[l + 1], 0, B, *self.ldims[l], F[l], st)
        lib.ssr_conv3d_wgrad_tc_up(self.vlow[l], F[l + 1], dy, dw, F[l] + F[l + 1], F[l], u['gscratch'], B,
                                   *self.ldims[l + 1], F[l], st)

    def _wgrad_async(self, *args, fn=None):
        """weight gradients only feed the optimiser, so they run on a side stream: the memory-bound elementwise kernels
        of the backward chain (BN / ELU / pooling gradients) then overlap with tensor-core-bound wgrad kernels instead
        of waiting for them.  Disabled while per-kernel profiling is on (timings would overlap)."""
        fn = fn or self._conv_wgrad
        if self.prof is not None or not self.overlap_wgrad:
            return fn(*args)
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
            self._ev_pool = [torch.cuda.Event() for _ in range(64)]
            self._ev_i = 0
        ev = self._ev_pool[self._ev_i % len(self._ev_pool)]
        self._ev_i += 1
        ev.record()                                   # the gradient this wgrad reads is complete on the main stream
        with torch.cuda.stream(self._side):
            self._side.wait_event(ev)
            fn(*args)
        self._side_busy = True

    def _wgrad_join(self):
        if self._side is not None and self._side_busy:
            ev = self._ev_pool[self._ev_i % len(self._ev_pool)]
            self._ev_i += 1
            ev.record(self._side)
            torch.cuda.current_stream().wait_event(ev)
            self._side_busy = False

    def _wgrad_tc_ok(self, c1, c2, cout):
        return getattr(self, 'wgrad_tc', False) and (c1 % 8 == 0) and (c2 % 8 == 0) and cout % 8 == 0

    def _packed_w(self, name, mode, c1, c2, cout, tag=0, src=None, buf=None):
        """packed (K-major, zero padded) copy of a kernel for the tcgen05 path; refreshed after every optimiser step.
        src: tensor to pack instead of the layer's kernel (derived kernels of the parity path); buf: destination view."""
        if self._packed_dirty:
            self._packed_valid = set()
            self._packed_dirty = False
            self._up_valid = False
        key = (name, mode, tag)
        if key not in self._packed:
            n = lib.ssr_conv3d_packed_size(c1, c2, cout, mode)
            self._packed[key] = buf if buf is not None else torch.empty(n, dtype=torch.float32, device=self.device)
            assert self._packed[key].numel() == n
            self._packed_args[key] = (c1, c2, cout)
            self._packed_src[key] = src if src is not None else self.p[name + '/kernel']
        if key not in self._packed_valid:
            if src is not None:
                self._up_weights_all(stream_ptr())
            lib.ssr_conv3d_pack_weights(self._packed_src[key], self._packed[key], c1, c2, cout, mode, stream_ptr())
            self._packed_valid.add(key)
        return self._packed[key]

    # ---- parity path of the decoder convolutions (conv3d_tc_up_kernel) -------------------------------------------------
    def _up_state(self, l):
        """scratch of decoder level l: skip part of the kernel, the 8 effective kernels of the upsampled part, and the
        buffers of their packed copies (8 parity classes back to back)."""
        if l not in self._up:
            F, dev = self.feats, self.device
            cs, cu, co = F[l], F[l + 1], F[l]
            nf = lib.ssr_conv3d_packed_size(cu, 0, co, 0)
            nd = lib.ssr_conv3d_packed_size(cu, 0, co, 1)
            self._up[l] = dict(wskip=torch.empty(27 * cs * co, dtype=torch.float32, device=dev),
                               weff=torch.empty(8 * 27 * cu * co, dtype=torch.float32, device=dev),
                               fwd8=torch.empty(8 * nf, dtype=torch.float32, device=dev), nf=nf,
                               dgr8=torch.empty(8 * nd, dtype=torch.float32, device=dev), nd=nd)
        return self._up[l]

    def _up_weights_all(self, st):
        if self._up_valid:
            return
        F, L = self.feats, self.L
        for l in self.up_levels:
            u = self._up_state(l)
            lib.ssr_conv3d_up_weights(self.p['unet_conv_uparm_%d_0/kernel' % (L + (L - 2 - l))], F[l], F[l + 1], F[l], u['wskip'],
                                      u['weff'], st)
            if self._up_k2n_ok(l):
                if 'wpk' not in u:
                    u['wpk'] = torch.empty(4 * 8 * 96 * 32, dtype=torch.float32, device=self.device)
                lib.ssr_conv3d_pack_up_k2n(u['weff'], u['wpk'], F[l + 1], st)
        self._up_valid = True

    def _up_k2n_ok(self, l):
        return self.up_k2n and self.fwd_k2n and self.feats[l] == 24 and self.feats[l + 1] <= 64

    def _up_packs(self, l, which):
        """packed weights of the parity kernels of level l: 'fwd8' (mode 0 per class) or 'dgr8' (mode 1 per class)."""
        F, L = self.feats, self.L
        u = self._up_state(l)
        name = 'unet_conv_uparm_%d_0' % (L + (L - 2 - l))
        cu, co = F[l + 1], F[l]
        if which == 'fwd8c' and which not in u:      # hi / lo packing of the 8 effective kernels (compensated forward)
            u['nfc'] = lib.ssr_conv3d_packed_size(cu, cu, co, 5)
            u[which] = torch.empty(8 * u['nfc'], dtype=torch.float32, device=self.device)
        if which == 'fwd8h' and which not in u:      # same for the hybrid scheme (TF32 hi chunks + bf16 chunks)
            u['nfh'] = lib.ssr_conv3d_packed_size(cu, cu, co, 7)
            u[which] = torch.empty(8 * u['nfh'], dtype=torch.float32, device=self.device)
        if which == 'fwd8b' and which not in u:      # bf16x3 scheme (w1 chunks + w2 chunks, all bf16)
            u['nfb'] = lib.ssr_conv3d_packed_size(cu, cu, co, 9)
            u[which] = torch.empty(8 * u['nfb'], dtype=torch.float32, device=self.device)
        n, mode, c2 = {'fwd8': (u['nf'], 0, 0), 'dgr8': (u['nd'], 1, 0), 'fwd8c': (u.get('nfc'), 5, cu),
                       'fwd8h': (u.get('nfh'), 7, cu), 'fwd8b': (u.get('nfb'), 9, cu)}[which]
        for par in range(8):
            self._packed_w(name, mode, cu, c2, co, tag=('up', par), src=u['weff'][par * 27 * cu * co:(par + 1) * 27 * cu * co],
                           buf=u[which][par * n:(par + 1) * n])
        return u[which]

    def _repack(self, st):
        """TF32-rounded, K-major packed copies of every kernel the tensor-core path has used so far, one launch."""
        self._packed_dirty = False
        self._up_valid = False
        self._up_weights_all(st)
        if getattr(self, '_pack_jobs_n', 0) != len(self._packed):   # device job table, rebuilt when a copy is added
            rows = []
            for key, buf in self._packed.items():
                name, mode = key[0], key[1]
                c1, c2, cout = self._packed_args[key]
                rows.append([self._packed_src[key].data_ptr(), buf.data_ptr(), c1, c2, cout, mode])
            self._pack_jobs = torch.tensor(rows, dtype=torch.int64).to(self.device)
            self._pack_jobs_n = len(rows)
        lib.ssr_conv3d_pack_weights_batch(self._pack_jobs, self._pack_jobs_n, st)
        self._packed_valid = set(self._packed.keys())

    # -------------------------------------------------------------------------------------------------------------
    def forward(self, image, training=True):
        """image: float32 cuda tensor [B, X, Y, Z, Cin] (contiguous) -> pred [B, X, Y, Z, nb_labels]."""
        st = stream_ptr()
        B, L, F = self.B, self.L, self.feats
        assert image.is_cuda and image.dtype == torch.float32 and image.is_contiguous()
        assert list(image.shape) == [B] + self.dims + [self.cin], image.shape
        self._image = image
        if self._packed_dirty and self._packed:          # refresh every packed kernel copy once per optimiser step
            self._repack(st)
        if self._pack_event is not None:                  # packing ran on the side stream right after the optimiser step
            torch.cuda.current_stream().wait_event(self._pack_event)
            self._pack_event = None
        x, cx = image, self.cin
        for l in range(L):
            if (l == 0 and cx <= 2 and self.k == 3 and F[0] in (8, 24) and
                    self._fuse_split_ok('unet_conv_downarm_0_1', F[0], F[0])):
                self._timed('fwd_ref', 0, cx, F[0], lambda: self._first_fwd_split(x, cx))
            else:
                self._conv_fwd('unet_conv_downarm_%d_0' % l, x, cx, None, 0, self.h0[l], l, F[l])
            bn = 'unet_bn_down_%d' % l
            fused = training and self._k2n_epi_ok(F[l], F[l], l, 'unet_conv_downarm_%d_1' % l)
            self._conv_fwd('unet_conv_downarm_%d_1' % l, self.h0[l], F[l], None, 0, self.h1[l], l, F[l],
                           stats_sums=self.sums if fused else None)
            self._bn_stats(bn, self.h1[l], self.nvox[l], F[l], self.stats_enc[l], training, have_sums=fused)
            if l < L - 1:
                lib.ssr_bn_apply(self.h1[l], self.inp[l + 1], self.stats_enc[l], B, *self.ldims[l], F[l], 1, 0, 0, st)
                x, cx = self.inp[l + 1], F[l]
        prev, prev_stats, prev_l = self.h1[L - 1], self.stats_enc[L - 1], L - 1
        for d in range(L - 1):
            l = L - 2 - d
            if l in self.up_levels:
                lib.ssr_bn_apply(prev, self.vlow[l], prev_stats, B, *self.ldims[prev_l], F[prev_l], 0, 0, 0, st)
                if training and not self.up_wgrad:      # the weight gradient still reads the upsampled tensor
                    lib.ssr_bn_apply(prev, self._u(l), prev_stats, B, *self.ldims[prev_l], F[prev_l], 2, 0, 0, st)
                self._conv_fwd_up('unet_conv_uparm_%d_0' % (L + d), l)
            else:
                lib.ssr_bn_apply(prev, self._u(l), prev_stats, B, *self.ldims[prev_l], F[prev_l], 2, 0, 0, st)
                self._conv_fwd('unet_conv_uparm_%d_0' % (L + d), self.h1[l], F[l], self.u[l], F[l + 1], self.g0[l], l, F[l])
            fused = training and self._k2n_epi_ok(F[l], F[l], l, 'unet_conv_uparm_%d_1' % (L + d))
            self._conv_fwd('unet_conv_uparm_%d_1' % (L + d), self.g0[l], F[l], None, 0, self.g1[l], l, F[l],
                           stats_sums=self.sums if fused else None)
            self._bn_stats('unet_bn_up_%d' % d, self.g1[l], self.nvox[l], F[l], self.stats_dec[l], training,
                           have_sums=fused)
            prev, prev_stats, prev_l = self.g1[l], self.stats_dec[l], l
        if L > 1 and not self.materialise_feat:
            # the last BatchNorm is folded into the head kernel (raw g1 + stats): no normalised feature tensor
            self._feat_src, self._feat_stats = self.g1[0], self.stats_dec[0]
            return self.g1[0]
        lib.ssr_bn_apply(self.g1[0], self.feat, self.stats_dec[0], B, *self.ldims[0], F[0], 0, 0, 0, st)
        self._feat_src, self._feat_stats = self.feat, None
        return self.feat

    def _first_fwd_split(self, x, cx):
        """first convolution (exact fp32 on the CUDA cores) + the bf16 split of its output for downarm_0_1's hybrid forward"""
        F0, name = self.feats[0], 'unet_conv_downarm_0_0'
        lib.ssr_conv3d_first_fwd_split(x, cx, self.p[name + '/kernel'], self.p[name + '/bias'], self.h0[0],
                                       self._lo_buf(self.nvox[0] * F0), self.B, *self.ldims[0], F0, 1, stream_ptr())
        self._lo_src = (self.h0[0].data_ptr(), self.nvox[0], F0)

    def _u(self, l):
        if self.u[l] is None:
            self.u[l] = torch.empty((self.nvox[l], self.feats[l + 1]), dtype=torch.float32, device=self.device)
        return self.u[l]

    def _bn_stats(self, bn, x, nvox, C, stats, training, have_sums=False):
        st = stream_ptr()
        if training and have_sums:       # self.sums was filled by the epilogue of the convolution that wrote x
            lib.ssr_bn_finalize(self.sums, nvox, C, self.p[bn + '/gamma'], self.p[bn + '/beta'],
                                self.moving[bn + '/moving_mean'], self.moving[bn + '/moving_variance'], BN_EPS,
                                self.bn_momentum, stats, st)
        elif training:
            lib.ssr_bn_stats(x, nvox, C, self.p[bn + '/gamma'], self.p[bn + '/beta'], self.moving[bn + '/moving_mean'],
                             self.moving[bn + '/moving_variance'], BN_EPS, self.bn_momentum, self.sums, stats, st)
        else:
            lib.ssr_bn_stats_inference(C, self.p[bn + '/gamma'], self.p[bn + '/beta'], self.moving[bn + '/moving_mean'],
                                       self.moving[bn + '/moving_variance'], BN_EPS, stats, st)

    def forward_frozen(self, image):
        """training-mode forward (batch statistics) that leaves the moving statistics untouched -> [B,X,Y,Z,nb_labels]:
        what Keras computes for this network inside a model that is being fitted while the network is `trainable = False`."""
        keep, self.bn_momentum = self.bn_momentum, 1.
        try:
            self.forward(image, training=True)
        finally:
            self.bn_momentum = keep
        zeros = torch.zeros((self.nvox[0], self.nb_labels), dtype=torch.float32, device=self.device)
        UNet3D._head(self, zeros, 'l1', None, None, train=False)
        return self.pred.view(self.B, *self.dims, self.nb_labels)

    def predict(self, image):
        """inference-mode forward (moving BN statistics) -> [B,X,Y,Z,nb_labels] tensor."""
        self.forward(image, training=False)
        zeros = torch.zeros((self.nvox[0], self.nb_labels), dtype=torch.float32, device=self.device)
        self._head(zeros, 'l1', None, None, train=False)
        return self.pred.view(self.B, *self.dims, self.nb_labels)

    def _head(self, target, metric, residual, loss_cropping, train):
        import ctypes
        st = stream_ptr()
        res_idx = crop_size = crop_begin = None                      # HOST int arrays (kept alive on self)
        if residual is not None:
            residual = [int(c) for c in residual]
            assert len(residual) == self.nb_labels and all(0 <= c < self.cin for c in residual)
            self._res_keep = (ctypes.c_int * 4)(*(residual + [0] * (4 - len(residual))))
            res_idx = ctypes.cast(self._res_keep, ctypes.c_void_p)
        if loss_cropping is not None:
            lc = [int(loss_cropping)] * 3 if isinstance(loss_cropping, (int, np.integer)) else [int(v) for v in loss_cropping]
            cb = [int((self.dims[i] - lc[i]) / 2) for i in range(3)]
            self._crop_keep = ((ctypes.c_int * 3)(*lc), (ctypes.c_int * 3)(*cb))
            crop_size = ctypes.cast(self._crop_keep[0], ctypes.c_void_p)
            crop_begin = ctypes.cast(self._crop_keep[1], ctypes.c_void_p)
        name = 'unet_likelihood'
        self._head_sums_valid = False
        if train and self.head_bn_sums and self._feat_stats is not None:
            # folded BatchNorm: the head also delivers the two reductions of that BatchNorm's backward
            C = self.feats[0]
            if getattr(self, '_xdot', None) is None:
                self._xdot = torch.empty(C * self.nb_labels, dtype=torch.float32, device=self.device)
                self._sums_head = torch.empty(2 * C, dtype=torch.float64, device=self.device)
            lib.ssr_head_loss_bnsums(self._feat_src, self._feat_stats, self.p[name + '/kernel'], self.p[name + '/bias'],
                                     self._image if residual is not None else None, self.cin, res_idx, target, self.pred,
                                     self.dbn_dec[0], self.g[name + '/kernel'], self.g[name + '/bias'], self.loss_buf,
                                     self.gout, self.B, *self.dims, C, self.nb_labels, 1 if metric == 'l1' else 2,
                                     crop_size, crop_begin, self._xdot, self._sums_head, st)
            self._head_sums_valid = True
            return
        lib.ssr_head_loss(self._feat_src, self._feat_stats, self.p[name + '/kernel'], self.p[name + '/bias'],
                          self._image if residual is not None else None, self.cin, res_idx, target, self.pred,
                          self.dbn_dec[0] if train else None, self.g[name + '/kernel'] if train else None,
                          self.g[name + '/bias'] if train else None, self.loss_buf, self.gout if train else None,
                          self.B, *self.dims, self.feats[0], self.nb_labels, 1 if metric == 'l1' else 2, crop_size,
                          crop_begin, 1 if train else 0, st)

    # -------------------------------------------------------------------------------------------------------------
    def loss_and_grad(self, image, target, metric='l1', work_with_residual_channel=None, loss_cropping=None):
        """forward (training BN) + loss + full backward.  Gradients are left in self.grads (flat).  Returns the loss
        as a 1-element float64 cuda tensor (no host sync)."""
        assert metric in ('l1', 'l2'), "regression_metric %r is out of scope of this build ('l1'/'l2' only)" % metric
        self._alloc_bwd()
        st = stream_ptr()
        B, L, F = self.B, self.L, self.feats
        assert target.is_cuda and target.dtype == torch.float32 and target.is_contiguous()
        self.forward(image, training=True)
        self.grads.zero_()
        self._head(target, metric, work_with_residual_channel, loss_cropping, train=True)
        # The backward chain (dgrad + elementwise gradients) is the critical path: it runs on a high-priority stream, the
        # weight gradients on a normal-priority side stream, so pending dgrad CTAs are dispatched ahead of wgrad CTAs
        # and wgrad fills the SMs while the chain is in its memory-bound elementwise phases.
        overlap = self.overlap_wgrad and self.prof is None
        cur = torch.cuda.current_stream()
        if overlap:
            if self._hp is None:
                self._hp = torch.cuda.Stream(device=self.device, priority=-1)
            self._hp.wait_stream(cur)
        with (torch.cuda.stream(self._hp) if overlap else contextlib.nullcontext()):
            st = stream_ptr()
            # ---- decoder, shallow to deep -----------------------------------------------------------------------
            for l in range(L - 1):
                d = L - 2 - l
                c0, c1n = 'unet_conv_uparm_%d_0' % (L + d), 'unet_conv_uparm_%d_1' % (L + d)
                bn = 'unet_bn_up_%d' % d
                if l == 0 and getattr(self, '_head_sums_valid', False):
                    lib.ssr_bn_bwd_sums(self.dbn_dec[l], self.g1[l], self.stats_dec[l], self.nvox[l], F[l], None, 0, 0, 1,
                                        self.ga[l], self.g[bn + '/gamma'], self.g[bn + '/beta'], self.g[c1n + '/bias'],
                                        self._sums_head, st)
                else:
                    lib.ssr_bn_bwd(self.dbn_dec[l], self.g1[l], self.stats_dec[l], self.nvox[l], F[l], None, 0, 0, 1,
                                   self.ga[l], self.g[bn + '/gamma'], self.g[bn + '/beta'], self.g[c1n + '/bias'], self.sums, st)
                self._wgrad_async(c1n, self.g0[l], F[l], None, 0, self.ga[l], l, F[l])
                if self._k2n_epi_ok(F[l], F[l], l):
                    self._conv_dgrad(c1n, self.ga[l], self.gb[l], l, F[l], F[l], elu_h=self.g0[l], dbias=self.g[c0 + '/bias'])
                else:
                    self._conv_dgrad(c1n, self.ga[l], self.gb[l], l, F[l], F[l])
                    lib.ssr_elu_bwd(self.gb[l], 0, 0, self.g0[l], None, self.nvox[l], F[l], self.gb[l], self.g[c0 + '/bias'], st)
                if l in self.up_levels and self.up_wgrad:
                    self._wgrad_async(c0, l, self.gb[l], fn=self._conv_wgrad_up)
                else:
                    self._wgrad_async(c0, self.h1[l], F[l], self.u[l], F[l + 1], self.gb[l], l, F[l])
                tgt = self.dbn_dec[l + 1] if l + 1 <= L - 2 else self.dbn_bott
                if l in self.up_levels:
                    self._conv_dgrad_up(c0, l, self.gb[l], tgt)
                else:
                    self._conv_dgrad(c0, self.gb[l], self.dcat[l], l, F[l] + F[l + 1], F[l])
                    lib.ssr_upsample_bwd(self.dcat[l], F[l] + F[l + 1], F[l], B, *self.ldims[l + 1], F[l + 1], tgt, st)
            # ---- encoder, deep to shallow -----------------------------------------------------------------------
            for l in range(L - 1, -1, -1):
                c0, c1n, bn = 'unet_conv_downarm_%d_0' % l, 'unet_conv_downarm_%d_1' % l, 'unet_bn_down_%d' % l
                if l == L - 1:
                    lib.ssr_bn_bwd(self.dbn_bott, self.h1[l], self.stats_enc[l], self.nvox[l], F[l], None, 0, 0, 1,
                                   self.ga_e[l], self.g[bn + '/gamma'], self.g[bn + '/beta'], self.g[c1n + '/bias'], self.sums, st)
                elif self.pool_bn_fusion and F[l] % 4 == 0 and 192 % (F[l] // 4) == 0:
                    add, add_stride = (self.dskip[l], F[l]) if l in self.up_levels else (self.dcat[l], F[l] + F[l + 1])
                    lib.ssr_pool_bn_bwd(self.dp[l + 1], self.h1[l], self.stats_enc[l], B, *self.ldims[l], F[l], add,
                                        add_stride, 0, 1, self.ga_e[l], self.g[bn + '/gamma'], self.g[bn + '/beta'],
                                        self.g[c1n + '/bias'], self.sums, st)
                else:
                    add, add_stride = (self.dskip[l], F[l]) if l in self.up_levels else (self.dcat[l], F[l] + F[l + 1])
                    lib.ssr_maxpool_bwd(self.dp[l + 1], self.h1[l], self.stats_enc[l], B, *self.ldims[l], F[l], self.ga_e[l],
                                        st)
                    lib.ssr_bn_bwd(self.ga_e[l], self.h1[l], self.stats_enc[l], self.nvox[l], F[l], add,
                                   add_stride, 0, 1, self.ga_e[l], self.g[bn + '/gamma'], self.g[bn + '/beta'],
                                   self.g[c1n + '/bias'], self.sums, st)
                self._wgrad_async(c1n, self.h0[l], F[l], None, 0, self.ga_e[l], l, F[l])
                if self._k2n_epi_ok(F[l], F[l], l):
                    self._conv_dgrad(c1n, self.ga_e[l], self.gb_e[l], l, F[l], F[l], elu_h=self.h0[l],
                                     dbias=self.g[c0 + '/bias'])
                else:
                    self._conv_dgrad(c1n, self.ga_e[l], self.gb_e[l], l, F[l], F[l])
                    lib.ssr_elu_bwd(self.gb_e[l], 0, 0, self.h0[l], None, self.nvox[l], F[l], self.gb_e[l],
                                    self.g[c0 + '/bias'], st)
                x, cx = (self._image, self.cin) if l == 0 else (self.inp[l], F[l - 1])
                self._wgrad_async(c0, x, cx, None, 0, self.gb_e[l], l, F[l])
                if self.grads_ready_hook is not None and l > 0:
                    self.grads_ready_hook(l)
                if l > 0:
                    self._conv_dgrad(c0, self.gb_e[l], self.dp[l], l, F[l - 1], F[l])
            self._wgrad_join()
        if overlap:
            cur.wait_stream(self._hp)
        return self.loss_buf

    def adam_step(self, lr=1e-4, lr_decay=0., beta1=.9, beta2=.999, eps=1e-7, grad_scale=1.):
        """keras.optimizers.Adam.get_updates (Keras 2.3.1) on the flat buffers, one launch."""
        lr_eff = lr
        if lr_decay > 0:
            lr_eff = lr * (1. / (1. + lr_decay * self.iterations))
        t = self.iterations + 1
        lr_t = lr_eff * (math.sqrt(1. - beta2 ** t) / (1. - beta1 ** t))
        lib.ssr_adam_flat(self.params, self.grads, self.adam_m, self.adam_v, self.n_params, lr_t, beta1, beta2, eps,
                          grad_scale, stream_ptr())
        self.iterations = t
        self._packed_dirty = True
        if self.overlap_wgrad and self._packed and self._side is not None:
            # re-pack on the side stream: overlaps with the next step's generator and first (exact fp32) convolution
            ev = self._ev_pool[self._ev_i % len(self._ev_pool)]
            self._ev_i += 1
            ev.record()
            with torch.cuda.stream(self._side):
                self._side.wait_event(ev)
                self._repack(stream_ptr())
                self._pack_event = torch.cuda.Event()
                self._pack_event.record()
