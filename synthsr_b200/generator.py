"""B200 generator engine: the graph of SynthSR/labels_to_image_model.py:32-266 as a fixed sequence of CUDA kernels.

`GeneratorPlan` holds the static bookkeeping (shapes, resolutions, blur sigmas) the reference computes at graph-build
time; `SynthGenerator.run` executes one step (labels, GMM parameters, draws) -> (image, target) on the current CUDA
stream through the C ABI in include/synthsr_b200.h.  No TensorFlow, no CPU fallback.
"""

import numpy as np
import torch

from . import draws as D
from ._lib import lib, stream_ptr

f32 = np.float32


def _as_list(v, n=None):
    if v is None:
        return None
    if isinstance(v, str):
        v = np.load(v)
    if isinstance(v, np.ndarray):
        v = np.squeeze(v).tolist()
    if isinstance(v, (bool, int, float, np.integer, np.floating, np.bool_)):
        v = [v]
    v = list(v)
    if n is not None:
        if len(v) == 1:
            v = v * n
        if len(v) != n:
            raise ValueError('expected a value of length 1 or %d, had %s' % (n, v))
    return v


def _res_array(v, n_channels):
    """utils.reformat_to_n_channels_array (ext/lab2im/utils.py:373-397)."""
    if v is None:
        return None
    if isinstance(v, str):
        v = np.load(v)
    v = np.array(v, dtype=np.float64)
    if v.ndim == 0:
        v = np.tile(v.reshape(1, 1), (n_channels, 3))
    elif v.ndim == 1:
        v = np.tile(v.reshape(1, 3), (n_channels, 1))
    return np.round(v.reshape(n_channels, 3), 3)


def _closest_div(n, m):
    return n if n % m == 0 else int(n / m) * m


def get_shapes(labels_shape, output_shape, atlas_res, target_res, padding_margin, output_div_by_n):
    """SynthSR/labels_to_image_model.py:269-335."""
    atlas_res = [float(v) for v in atlas_res]
    target_res = [float(v) for v in target_res]
    labels_shape = list(labels_shape)
    if padding_margin is not None:
        padding_margin = [int(v) for v in _as_list(padding_margin, 3)]
        labels_shape = [labels_shape[i] + 2 * padding_margin[i] for i in range(3)]
    factor = [atlas_res[i] / target_res[i] for i in range(3)] if atlas_res != target_res else None
    if output_shape is not None:
        output_shape = [int(v) for v in _as_list(output_shape, 3)]
        if factor is not None:
            output_shape = [min(int(labels_shape[i] * factor[i]), output_shape[i]) for i in range(3)]
        else:
            output_shape = [min(labels_shape[i], output_shape[i]) for i in range(3)]
        if output_div_by_n is not None:
            output_shape = [_closest_div(s, output_div_by_n) for s in output_shape]
        crop = [int(np.around(output_shape[i] / factor[i], 0)) for i in range(3)] if factor is not None else output_shape
    elif output_div_by_n is not None:
        if factor is not None:
            output_shape = [_closest_div(int(labels_shape[i] * factor[i]), output_div_by_n) for i in range(3)]
            crop = [int(np.around(output_shape[i] / factor[i], 0)) for i in range(3)]
        else:
            crop = [_closest_div(s, output_div_by_n) for s in labels_shape]
            output_shape = crop
    else:
        crop = labels_shape
        output_shape = [int(crop[i] * factor[i]) for i in range(3)] if factor is not None else crop
    return list(crop), list(output_shape), padding_margin


def blurring_sigma(current_res, down_res, mult_coef=None, thickness=None):
    """ext/lab2im/edit_tensors.py:41-83."""
    current_res = np.array(current_res, dtype=np.float64)
    down_res = np.array(down_res, dtype=np.float64)
    if thickness is not None:
        down_res = np.minimum(down_res, np.array(thickness, dtype=np.float64))
    if mult_coef is None:
        sigma = 0.75 * down_res / current_res
        sigma[down_res == current_res] = 0.5
    else:
        sigma = mult_coef * down_res / current_res
    sigma[down_res == 0] = 0
    return sigma


def gaussian_kernel(sigma, blur_mult=None):
    """Dense (||sigma|| <= 5) or separable kernels (ext/lab2im/edit_tensors.py:86-181, layers.py:720), float32.
    Returns a list of dense 3-D arrays to apply in sequence (one for dense, up to three for separable)."""
    sigma = [float(s) for s in sigma]
    max_sigma = np.array(sigma, dtype=np.float64)
    sig = np.array(sigma, dtype=f32)
    if blur_mult is not None:
        sig = (sig * np.asarray(blur_mult, dtype=f32)).astype(f32)
    ws = np.int32(np.ceil(2.5 * max_sigma) / 2) * 2 + 1
    c = f32(np.sqrt(2 * np.pi))
    if np.linalg.norm(max_sigma) > 5:
        out = []
        for i, w in enumerate(ws):
            if w > 1:
                loc = (np.arange(w).astype(f32) - f32((w - 1) / 2)).astype(f32)
                g = np.exp((-np.square(loc) / (f32(2) * sig[i] ** 2)).astype(f32) - np.log(c * sig[i]).astype(f32)).astype(f32)
                g = (g / np.sum(g, dtype=f32)).astype(f32)
                shape = [1, 1, 1]
                shape[i] = int(w)
                out.append(g.reshape(shape))
        return out
    if not any(sigma):
        return []
    mesh = np.meshgrid(*[np.arange(w) for w in ws], indexing='ij')
    diff = np.stack([(mesh[d].astype(f32) - f32((ws[d] - 1) / 2)).astype(f32) for d in range(3)], -1)
    is0 = sig == 0
    s1 = np.where(is0, f32(1), sig).astype(f32)
    exp_term = (-np.square(diff) / (f32(2) * s1 ** 2).astype(f32)).astype(f32)
    logt = np.log(np.where(is0, f32(1), (c * sig).astype(f32))).astype(f32)
    k = np.exp(np.sum((exp_term - logt).astype(f32), -1, dtype=f32)).astype(f32)
    return [(k / np.sum(k, dtype=f32)).astype(f32)]


def reliability_factors(resample_shape, downsample_shape):
    """per-axis tent weights of ext/lab2im/edit_tensors.py:313-329 (float64)."""
    up = np.array(resample_shape) / np.array(downsample_shape)
    out = []
    for i in range(3):
        loc_float = np.arange(0, resample_shape[i], up[i])
        loc_floor = np.int32(np.floor(loc_float))
        loc_ceil = np.int32(np.clip(loc_floor + 1, 0, resample_shape[i] - 1))
        tmp = np.zeros(resample_shape[i])
        tmp[loc_floor] = 1 - (loc_float - loc_floor)
        tmp[loc_ceil] = tmp[loc_ceil] + (loc_float - loc_floor)
        out.append(tmp)
    return out


def dynamic_sigma(atlas_res, resolution, thickness, mult_coef=.42):
    """blurring_sigma_for_downsampling, tensor branch (ext/lab2im/edit_tensors.py:66-81): [B,3] float32."""
    res = np.asarray(resolution, dtype=f32)
    down = np.minimum(res, np.asarray(thickness, dtype=f32)).astype(f32)
    sigma = ((f32(mult_coef) * down).astype(f32) / np.asarray(atlas_res, dtype=f32)).astype(f32)
    return np.where(down == 0, f32(0), sigma).astype(f32)


def dynamic_kernels(sigma, window, blur_mult=None):
    """per-example 1-D Gaussian kernels of DynamicGaussianBlur (ext/lab2im/edit_tensors.py:126-154 with a [B,3] sigma
    tensor): jitter per (example, axis), and -- like the reference's `g / tf.reduce_sum(g)` on the [B, window] tensor --
    normalisation by the sum over the whole batch.  Returns three float32 arrays [B, window[i]] (None for window 1)."""
    sig = np.asarray(sigma, dtype=f32)
    if blur_mult is not None:
        sig = (sig * np.asarray(blur_mult, dtype=f32)).astype(f32)
    out = []
    for i, w in enumerate(window):
        if w > 1:
            loc = (np.arange(w).astype(f32) - f32((w - 1) / 2)).astype(f32)[None, :]
            si = sig[:, i:i + 1]
            exp_term = (-np.square(loc) / (f32(2) * si ** 2).astype(f32)).astype(f32)
            g = np.exp(exp_term - np.log((f32(np.sqrt(2 * np.pi)) * si).astype(f32)).astype(f32)).astype(f32)
            out.append((g / np.sum(g, dtype=f32)).astype(f32))
        else:
            out.append(None)
    return out


def mimic_zooms(inshape, volume_res, subsample_res, resample_shape):
    """MimicAcquisition zoom factors for one example (ext/lab2im/layers.py:935-938) -> float32 [9] =
    down_zoom | up_zoom | resolution, the parameter block of ssr_mimic_acquisition."""
    full = (np.array(inshape) * np.array(volume_res, dtype=np.float64)).astype(f32)
    res = np.asarray(subsample_res, dtype=f32)
    down_shape = (full / res).astype(np.int32)
    down_zoom = (down_shape / np.array(inshape)).astype(f32)
    up_zoom = (np.array(resample_shape, dtype=np.int32) / down_shape).astype(f32)
    return np.concatenate([down_zoom, up_zoom, res]).astype(f32)


class GeneratorPlan:
    """Static configuration of one labels_to_image_model instance (same keyword names as the reference)."""

    def __init__(self, labels_shape, input_channels, output_channel, generation_labels, n_neutral_labels, atlas_res,
                 target_res, output_shape=None, output_div_by_n=None, padding_margin=None, flipping=True, aff=None,
                 scaling_bounds=0.15, rotation_bounds=15, shearing_bounds=0.012, translation_bounds=False,
                 nonlin_std=3., nonlin_shape_factor=.0625, simulate_registration_error=True, randomise_res=False,
                 data_res=None, thickness=None, downsample=False, build_reliability_maps=False, blur_range=1.15,
                 bias_field_std=.3, bias_shape_factor=.025):
        ic = input_channels
        self.input_channels = [bool(ic)] if isinstance(ic, (bool, int, np.bool_)) else [bool(v) for v in np.asarray(ic).ravel()]
        self.n_channels = len(self.input_channels)
        if output_channel is not None and not isinstance(output_channel, (list, tuple, np.ndarray)):
            output_channel = [int(output_channel)]
        self.output_channel = None if output_channel is None else [int(c) for c in output_channel]
        self.use_real_image = self.output_channel is None
        self.idx_first_input_channel = int(np.argmax(self.input_channels))
        sr = simulate_registration_error
        self.sim_reg = [bool(sr)] * self.n_channels if isinstance(sr, (bool, int, np.bool_)) else [bool(v) for v in sr]
        if isinstance(randomise_res, (bool, np.bool_)) or randomise_res is None:
            randomise_res = [bool(randomise_res)] * self.n_channels
        self.randomise_res = [bool(v) for v in randomise_res]
        assert len(self.randomise_res) == self.n_channels, 'randomise_res must have one entry per channel'
        self.labels_shape = [int(s) for s in labels_shape]
        atlas = _res_array(atlas_res, self.n_channels)
        data_res = None if data_res is None else (np.load(data_res) if isinstance(data_res, str) else data_res)
        thickness = None if thickness is None else (np.load(thickness) if isinstance(thickness, str) else thickness)
        if self.output_channel is not None:
            for idx in self.output_channel:
                if not self.input_channels[idx]:                       # labels_to_image_model.py:85-89
                    data_res = np.insert(np.array(data_res, dtype=np.float64).reshape(-1, 3), idx, 1, axis=0)
                    thickness = np.insert(np.array(thickness, dtype=np.float64).reshape(-1, 3), idx, 1, axis=0)
        self.data_res = atlas if data_res is None else _res_array(data_res, self.n_channels)
        self.thickness = self.data_res if thickness is None else _res_array(thickness, self.n_channels)
        if downsample:
            self.downsample = _as_list(downsample, self.n_channels)
        else:
            self.downsample = list(np.min(self.thickness - self.data_res, 1) < 0)
        self.atlas_res = atlas[0]
        self.target_res = self.atlas_res if target_res is None else _res_array(target_res, 1)[0]
        self.crop_shape, self.output_shape, self.padding_margin = get_shapes(
            self.labels_shape, output_shape, self.atlas_res, self.target_res, padding_margin, output_div_by_n)
        pm = self.padding_margin or [0, 0, 0]
        self.pad = [int(p) for p in pm]
        self.grid_shape = [self.labels_shape[i] + 2 * self.pad[i] for i in range(3)]
        self.generation_labels = np.asarray(generation_labels).astype(np.int64)
        self.n_neutral_labels = len(self.generation_labels) if n_neutral_labels is None else int(n_neutral_labels)
        self.flipping = bool(flipping)
        if self.flipping and aff is not None:
            # the reference flips along get_ras_axes(aff)[0] (labels_to_image_model.py:159-162); BrainGenerator always hands
            # np.eye(4) because label maps are re-oriented at load time.  The fused deformation kernel flips axis 0 only.
            from ext.lab2im.edit_volumes import get_ras_axes
            if int(get_ras_axes(np.asarray(aff, dtype=np.float64), 3)[0]) != 0:
                raise NotImplementedError('right/left flipping along axis %d: only volumes whose first axis is the R/L axis '
                                          '(aff aligned like np.eye(4), what BrainGenerator passes) are supported'
                                          % int(get_ras_axes(np.asarray(aff, dtype=np.float64), 3)[0]))
        self.scaling_bounds, self.rotation_bounds = scaling_bounds, rotation_bounds
        self.shearing_bounds, self.translation_bounds = shearing_bounds, translation_bounds
        self.apply_affine = any(b is not False for b in (scaling_bounds, rotation_bounds, shearing_bounds,
                                                         translation_bounds))
        self.nonlin_std = float(nonlin_std) if nonlin_std else 0.
        self.nonlin_shape_factor = nonlin_shape_factor
        if self.nonlin_std > 0:
            self.svf_small_shape = D.resample_shape(self.grid_shape, nonlin_shape_factor)          # layers.py:151
            self.svf_half_shape = [max(int(self.grid_shape[i] / 2), self.svf_small_shape[i]) for i in range(3)]
        else:
            self.svf_small_shape = self.svf_half_shape = None
        self.blur_range = blur_range
        self.build_reliability_maps = bool(build_reliability_maps)
        self.bias_field_std = float(bias_field_std) if bias_field_std else 0.
        self.bias_small_shape = D.resample_shape(self.crop_shape, bias_shape_factor)              # layers.py:1059
        # swap LUT for right/left flipping (layers.py:375-386)
        n_lab = len(self.generation_labels)
        self.swap_lut = None
        if self.flipping and self.n_neutral_labels != n_lab:
            split = np.split(self.generation_labels, [self.n_neutral_labels,
                                                      self.n_neutral_labels + int((n_lab - self.n_neutral_labels) / 2)])
            dest = np.concatenate((split[0], split[2], split[1]))
            lut = np.zeros(int(np.max(self.generation_labels)) + 1, dtype=np.int32)
            for s, t in zip(self.generation_labels, dest):
                lut[s] = t
            self.swap_lut = lut
        self.lut_len = int(np.max(self.generation_labels)) + 1
        # channel bookkeeping
        self.n_image_channels = sum(self.input_channels) * (2 if self.build_reliability_maps else 1)
        self.n_target_channels = 1 if self.use_real_image else len(self.output_channel)
        self.target_sigma = None
        if self.crop_shape != self.output_shape:
            self.target_sigma = blurring_sigma(self.atlas_res, self.target_res)
        self.acq_sigma = [blurring_sigma(self.atlas_res, self.data_res[i], .42, self.thickness[i])
                          for i in range(self.n_channels)]
        # randomise_res branch (labels_to_image_model.py:215-220): DynamicGaussianBlur(0.75 * max_res / atlas_res) with
        # max_res = 9 mm -> separable 1-D kernels of a fixed window (ext/lab2im/layers.py:808, edit_tensors.py:124)
        self.dyn_max_sigma = 0.75 * 9. / np.asarray(self.atlas_res, dtype=np.float64)
        self.dyn_window = [int(w) for w in (np.int32(np.ceil(2.5 * self.dyn_max_sigma) / 2) * 2 + 1)]
        # grid each input channel's acquisition chain (:199-238) runs on.  The reference REBINDS `channel` to the blurred,
        # resampled regression target at :193-195, so a channel that is both a target and an input continues on the OUTPUT
        # grid when target_res != atlas_res (pinned by executing the reference graph: tests/golden/reference_model.npz, B/E)
        self.chan_grid = []
        for i in range(self.n_channels):
            rebound = (not self.use_real_image) and i in self.output_channel and self.crop_shape != self.output_shape
            self.chan_grid.append(list(self.output_shape) if rebound else list(self.crop_shape))
        self.down_shape = []
        for i in range(self.n_channels):
            ds = list(self.chan_grid[i])                                       # edit_tensors.py:292-299 (tensor shape)
            if self.downsample[i] and list(self.data_res[i]) != list(self.atlas_res):
                ds = [int(self.chan_grid[i][k] * float(self.atlas_res[k]) / float(self.data_res[i][k])) for k in range(3)]
            self.down_shape.append(ds)

    @property
    def image_shape(self):
        return list(self.output_shape) + [self.n_image_channels]

    @property
    def target_shape(self):
        return list(self.output_shape) + [self.n_target_channels]


def gmm_luts(means, stds, generation_labels, lut_len):
    """label -> (mean, std) look-up tables [B, lut_len] of one channel, as SampleConditionalGMM builds them
    (ext/lab2im/layers.py:480-497): unlisted label values get 0.  Reference behaviour kept on purpose: the layer scatters
    the means of ALL batch elements into ONE table with tf.scatter_nd, which ADDS duplicate indices, and tiles that table
    over the batch -- for batchsize > 1 every example is sampled from the SUM over the batch of the drawn parameters
    (reproduced by executing the reference layer, tests/golden/make_reference_layer_goldens.py).  With batchsize 1, the
    reference's default and what every data-parallel rank runs, this is the plain per-example table."""
    means, stds = np.asarray(means, dtype=f32), np.asarray(stds, dtype=f32)
    B = means.shape[0]
    msum, ssum = means[0].copy(), stds[0].copy()
    for b in range(1, B):                                  # float32 accumulation in scatter order
        msum = (msum + means[b]).astype(f32)
        ssum = (ssum + stds[b]).astype(f32)
    ml = np.zeros((B, lut_len), dtype=f32)
    sl = np.zeros((B, lut_len), dtype=f32)
    ml[:, generation_labels] = msum[None]
    sl[:, generation_labels] = ssum[None]
    return ml, sl


class _Staging:
    """Pinned host buffers + one device buffer: all small per-step inputs go to the GPU in a single async copy.
    The host side is a ring of NBUF pinned buffers, each guarded by the event of its last copy: the host may enqueue
    steps ahead of the GPU, and must not overwrite a buffer whose copy has not executed yet."""
    NBUF = 4

    def __init__(self, nbytes, device):
        cuda = torch.cuda.is_available()
        self.hosts = [torch.empty(nbytes, dtype=torch.uint8, pin_memory=cuda) for _ in range(self.NBUF)]
        self.nps = [h.numpy() for h in self.hosts]
        self.events = [None] * self.NBUF
        self.cur = -1
        self.dev = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.host, self.np = self.hosts[0], self.nps[0]
        self.off = 0
        self.bytes_moved = 0

    def reset(self):
        self.cur = (self.cur + 1) % self.NBUF
        if self.events[self.cur] is not None:
            self.events[self.cur].synchronize()          # the copy that last read this host buffer has completed
        self.host, self.np = self.hosts[self.cur], self.nps[self.cur]
        self.off = 0

    def put(self, arr):
        arr = np.ascontiguousarray(arr)
        nb = arr.nbytes
        start = (self.off + 15) // 16 * 16
        if start + nb > self.np.size:
            raise RuntimeError('staging buffer too small')
        self.np[start:start + nb] = arr.view(np.uint8).reshape(-1)
        self.off = start + nb
        return self.dev.data_ptr() + start

    def flush(self):
        if self.off:
            self.dev[:self.off].copy_(self.host[:self.off], non_blocking=True)
            self.bytes_moved += self.off
            if self.dev.is_cuda:
                if self.events[self.cur] is None:
                    self.events[self.cur] = torch.cuda.Event()
                self.events[self.cur].record()


class SynthGenerator:
    def __init__(self, plan, batchsize=1, device='cuda'):
        self.plan = plan
        self.B = int(batchsize)
        self.device = torch.device(device)
        if self.device.type != 'cuda' and not getattr(lib, 'host_emulation', False):
            raise RuntimeError('SynthGenerator needs a CUDA device: there is no CPU path (the CPU suite swaps `lib` for '
                               'tests/host_emulator.py to exercise the orchestration only)')
        p = plan
        B = self.B
        dev = self.device
        nc = int(np.prod(p.crop_shape))
        nt = max(nc, int(np.prod(p.output_shape)))
        self.labels = torch.empty((B, *p.crop_shape), dtype=torch.int32, device=dev)
        self.raw = torch.empty((B, nc), dtype=torch.float32, device=dev)
        self.tmp_a = torch.empty((B, nt), dtype=torch.float32, device=dev)
        self.tmp_b = torch.empty((B, nt), dtype=torch.float32, device=dev)
        self.tmp_c = torch.empty((B, nt), dtype=torch.float32, device=dev)
        self.minmax = torch.empty((B, 2), dtype=torch.int32, device=dev)
        self.image = torch.empty((B, *p.output_shape, p.n_image_channels), dtype=torch.float32, device=dev)
        self.target = torch.empty((B, *p.output_shape, p.n_target_channels), dtype=torch.float32, device=dev)
        if p.nonlin_std > 0:
            self.svf_half = torch.empty((B, *p.svf_half_shape, 3), dtype=torch.float32, device=dev)
            self.svf_tmp = torch.empty_like(self.svf_half)
        if p.use_real_image:
            self.real = torch.empty((B, nc), dtype=torch.float32, device=dev)
        nd = max(int(np.prod(s)) for s in p.down_shape)
        if any(p.randomise_res):
            nd = max(nd, nt)                 # also holds the acquisition distance map when it has to be warped
        self.down = torch.empty((B, nd), dtype=torch.float32, device=dev)
        small = 4 * B * (16 + 3 * int(np.prod(p.svf_small_shape or [1])) + 3 + 1 + 2 * p.lut_len * p.n_channels
                         + p.n_channels * (int(np.prod(p.bias_small_shape)) + 2048 + 64)) + 8 * sum(p.output_shape) * 2
        self.stage = _Staging(small + (1 << 16), dev)
        self.philox_step = 0

    # -----------------------------------------------------------------------------------------------------------
    def run(self, labels, means, stds, draws, real_image=None, seed=0, keep=None):
        """labels: int32 cuda tensor [B, *labels_shape]; means/stds: array-like [B, L, C]; draws: see draws.py.
        Returns (image [B,*out,Cimg], target [B,*out,Cout]) float32 cuda tensors (views of internal buffers).
        `keep`: optional dict that receives clones of intermediates (tests)."""
        p, B = self.plan, self.B
        st = stream_ptr()
        sg = self.stage
        sg.reset()
        assert labels.dtype == torch.int32 and labels.device.type == self.device.type and labels.is_contiguous()
        assert list(labels.shape) == [B] + p.labels_shape, (labels.shape, p.labels_shape)
        means = np.asarray(means, dtype=f32).reshape(B, len(p.generation_labels), p.n_channels)
        stds = np.asarray(stds, dtype=f32).reshape(B, len(p.generation_labels), p.n_channels)

        # ---- small host->device inputs, one copy -----------------------------------------------------------
        aff_ptr = None
        if p.apply_affine:
            aff = np.stack([D.build_affine(
                draws['aff_rotation'][b] if draws.get('aff_rotation') is not None else None,
                draws['aff_shearing'][b] if draws.get('aff_shearing') is not None else None,
                draws['aff_scaling'][b] if draws.get('aff_scaling') is not None else None,
                draws['aff_translation'][b] if draws.get('aff_translation') is not None else None) for b in range(B)])
            aff_ptr = sg.put(aff)
            if keep is not None:
                keep['affine'] = aff
        svf_ptr = None
        if p.nonlin_std > 0:
            svf = (np.asarray(draws['svf_normal'], dtype=f32) * f32(draws['svf_std'])).astype(f32)
            svf_ptr = sg.put(svf)
        crop_ptr = sg.put(np.asarray(draws['crop_idx'], dtype=np.int32)) if p.crop_shape != p.grid_shape else None
        flip_ptr = sg.put(np.asarray(draws['flip'], dtype=np.uint8)) if p.flipping else None
        lut_ptr = sg.put(p.swap_lut) if p.swap_lut is not None else None
        chan = []
        for i in range(p.n_channels):
            c = {}
            ml, sl = gmm_luts(means[:, :, i], stds[:, :, i], p.generation_labels, p.lut_len)
            c['mean'], c['std'] = sg.put(ml), sg.put(sl)
            c['bias'] = None
            c['apply'] = 0
            if p.input_channels[i] and p.bias_field_std > 0:
                bs = (np.asarray(draws['bias_normal_%d' % i], dtype=f32) *
                      np.asarray(draws['bias_std_%d' % i], dtype=f32).reshape(B, 1, 1, 1)).astype(f32)
                c['apply'] = int(bool(draws['bias_apply_%d' % i]))
                if c['apply']:
                    c['bias'] = sg.put(bs)
            gam = (np.asarray(draws['gamma_normal_%d' % i], dtype=f32) * f32(.5)).astype(f32)
            c['gamma'] = sg.put(np.exp(gam).astype(f32))
            c['k05'] = sg.put(gaussian_kernel([.5, .5, .5])[0])
            if p.input_channels[i] and p.randomise_res[i]:                    # labels_to_image_model.py:215-220
                jit = p.blur_range is not None and p.blur_range != 1
                sig = dynamic_sigma(p.atlas_res, draws['res_%d' % i], draws['thick_%d' % i], .42)
                ks = dynamic_kernels(sig, p.dyn_window, draws['blur_mult_dyn_%d' % i] if jit else None)
                c['kdyn'] = [None if k is None else sg.put(k) for k in ks]   # [B, window] per axis
                c['mimic'] = sg.put(np.stack([mimic_zooms(p.chan_grid[i], p.atlas_res, draws['res_%d' % i][b],
                                                          p.output_shape) for b in range(B)]))
                c['kacq'] = []
            elif p.input_channels[i]:
                mult = draws.get('blur_mult_%d' % i) if (p.blur_range is not None and p.blur_range != 1) else None
                ks = gaussian_kernel(list(p.acq_sigma[i]), mult)
                c['kacq'] = [(sg.put(k), k.shape) for k in ks]
            if p.input_channels[i]:
                do_reg = p.sim_reg[i] and i != p.idx_first_input_channel
                c['reg'] = do_reg
                if do_reg:
                    T = [D.build_affine(rotation=draws['reg_rot_%d' % i][b], translation=draws['reg_trans_%d' % i][b])
                         for b in range(B)]
                    Tinv = [np.linalg.inv(t.astype(np.float64)).astype(f32) for t in T]
                    Terr = [D.build_affine(rotation=draws['reg_err_rot_%d' % i][b],
                                           translation=draws['reg_err_trans_%d' % i][b]) for b in range(B)]
                    c['T'] = sg.put(np.stack(T))
                    c['Tie'] = sg.put(np.stack([D.matmul4(Terr[b], Tinv[b]) for b in range(B)]))
                if p.build_reliability_maps and p.down_shape[i] != p.chan_grid[i] and not p.randomise_res[i]:
                    c['rel'] = [sg.put(f) for f in reliability_factors(p.output_shape, p.down_shape[i])]
            if (not p.use_real_image) and i in p.output_channel and p.target_sigma is not None:
                c['ktgt'] = [(sg.put(k), k.shape) for k in gaussian_kernel(list(p.target_sigma))]
            chan.append(c)
        if p.use_real_image and p.target_sigma is not None:
            ktgt_real = [(sg.put(k), k.shape) for k in gaussian_kernel(list(p.target_sigma))]
        sg.flush()

        g, pad, cs = p.grid_shape, p.pad, p.crop_shape
        # ---- deformation field: small SVF -> half res -> integrate (layers.py:188-195) -----------------------
        h = [0, 0, 0]
        field_ptr = None
        if p.nonlin_std > 0:
            h = p.svf_half_shape
            lib.ssr_resize(svf_ptr, self.svf_half, B, *p.svf_small_shape, *h, 3, 0, 0, 0, st)
            lib.ssr_svf_integrate(self.svf_half, self.svf_tmp, B, *h, 7, st)
            field_ptr = self.svf_half.data_ptr()
            if keep is not None:
                keep['integrated'] = self.svf_half.clone()
        # ---- labels: pad + full-res field + affine + nearest + crop + flip + swap, one kernel ------------------
        lib.ssr_deform_labels_nearest(labels, self.labels, aff_ptr, field_ptr, B, *g, *pad, *h, crop_ptr, *cs,
                                      flip_ptr, lut_ptr, p.lut_len if lut_ptr else 0, st)
        if keep is not None:
            keep['labels'] = self.labels.clone()
        if p.use_real_image:
            assert real_image is not None and real_image.dtype == torch.float32 and real_image.device.type == self.device.type
            lib.ssr_warp_linear(real_image.contiguous(), self.real, aff_ptr, field_ptr, B, *g, *pad, *h, crop_ptr, *cs,
                                flip_ptr, st)
        # ---- per-channel chain --------------------------------------------------------------------------------
        noise = draws.get('gmm_normal')
        noise_t = None
        if noise is not None:
            noise_t = torch.from_numpy(np.ascontiguousarray(np.moveaxis(np.asarray(noise, dtype=f32), -1, 0))).to(
                self.device, non_blocking=True)                                   # [C, B, X, Y, Z]
        self.philox_step += 1
        out_c = 0
        tgt_c = 0
        os_ = p.output_shape
        for i, c in enumerate(chan):
            cs = p.crop_shape                                                  # GMM / bias / intensity / blur(.5): crop grid
            nptr = noise_t[i].data_ptr() if noise_t is not None else None
            bsh = p.bias_small_shape if c['bias'] else [0, 0, 0]
            lib.ssr_gmm_bias_minmax(self.labels, c['mean'], c['std'], p.lut_len, nptr, int(seed),
                                    (self.philox_step << 8) + (i << 4), c['bias'], *bsh, c['apply'], 300., self.raw,
                                    self.minmax, B, *cs, st)
            if keep is not None:
                keep['raw_%d' % i] = self.raw.clone()
            # normalise + gamma + GaussianBlur(.5)  (labels_to_image_model.py:184-186)
            lib.ssr_blur3d(self.raw, self.tmp_a, c['k05'], 3, 3, 3, self.minmax, c['gamma'], B, *cs, 1, 0, 1, 0, st)
            if keep is not None:
                keep['blur_%d' % i] = self.tmp_a.clone()
            cur = self.tmp_a
            if (not p.use_real_image) and i in p.output_channel:
                rebound = p.input_channels[i] and p.chan_grid[i] != cs         # :193-195 `channel` now IS the target
                for _ in range(p.output_channel.count(i)):
                    res = self._emit_target(self.tmp_a, c.get('ktgt'), tgt_c, st, rebound)
                    tgt_c += 1
                if rebound:
                    cur = res
            if not p.input_channels[i]:
                continue
            bufs = [self.tmp_a, self.tmp_b, self.tmp_c]

            def free(*used):
                return next(t for t in bufs if all(t is not u for u in used))

            cs = p.chan_grid[i]                                                # grid of this channel's acquisition chain
            if c['reg']:                                                       # :202-208
                dst = free(cur)
                lib.ssr_warp_linear(cur, dst, c['T'], None, B, *cs, 0, 0, 0, 0, 0, 0, None, *cs, None, st)
                cur = dst
            if p.randomise_res[i]:                                             # :215-220
                nv_in = int(np.prod(cs))
                for ax, kp in enumerate(c['kdyn']):                            # DynamicGaussianBlur: per-example kernels
                    if kp is None:
                        continue
                    ksh = [1, 1, 1]
                    ksh[ax] = p.dyn_window[ax]
                    dst = free(cur)
                    for b in range(B):
                        lib.ssr_blur3d(cur.data_ptr() + 4 * b * nv_in, dst.data_ptr() + 4 * b * nv_in,
                                       kp + 4 * b * p.dyn_window[ax], *ksh, None, None, 1, *cs, 1, 0, 1, 0, st)
                    cur = dst
                want_rel = p.build_reliability_maps
                if c['reg']:
                    dst, dd = free(cur), self.down
                    lib.ssr_mimic_acquisition(cur, dst, dd if want_rel else None, c['mimic'], B, *cs, *os_, 1, 0, 1, 0, st)
                    d2 = free(dst)
                    lib.ssr_warp_linear(dst, d2, c['Tie'], None, B, *os_, 0, 0, 0, 0, 0, 0, None, *os_, None, st)
                    lib.ssr_copy_strided(d2, self.image, B * int(np.prod(os_)), 1, 0, p.n_image_channels, out_c, st)
                    out_c += 1
                    if want_rel:
                        lib.ssr_warp_linear(dd, d2, c['Tie'], None, B, *os_, 0, 0, 0, 0, 0, 0, None, *os_, None, st)
                        lib.ssr_copy_strided(d2, self.image, B * int(np.prod(os_)), 1, 0, p.n_image_channels, out_c, st)
                        out_c += 1
                else:
                    lib.ssr_mimic_acquisition(cur, self.image, self.image if want_rel else None, c['mimic'], B, *cs, *os_,
                                              p.n_image_channels, out_c, p.n_image_channels, out_c + 1, st)
                    out_c += 2 if want_rel else 1
                continue
            ksteps = c['kacq']                                                 # :223-224
            direct = (not c['reg']) and p.down_shape[i] == cs and os_ == cs
            for n, (kp, ksh) in enumerate(ksteps):
                if direct and n == len(ksteps) - 1:
                    lib.ssr_blur3d(cur, self.image, kp, *ksh, None, None, B, *cs, 1, 0, p.n_image_channels, out_c, st)
                    cur = None
                else:
                    dst = free(cur)
                    lib.ssr_blur3d(cur, dst, kp, *ksh, None, None, B, *cs, 1, 0, 1, 0, st)
                    cur = dst
            if cur is not None:
                src, sshape = cur, cs
                if p.down_shape[i] != cs:                                      # edit_tensors.py:295-299 nearest down
                    lib.ssr_resize(src, self.down, B, *cs, *p.down_shape[i], 1, 1, 0, 0, st)
                    src, sshape = self.down, p.down_shape[i]
                if os_ != sshape:                                              # edit_tensors.py:302-304 linear up
                    if c['reg']:
                        dst = free(src)
                        lib.ssr_resize(src, dst, B, *sshape, *os_, 1, 0, 0, 0, st)
                        src = dst
                    else:
                        lib.ssr_resize(src, self.image, B, *sshape, *os_, 1, 0, p.n_image_channels, out_c, st)
                        src = None
                if c['reg']:                                                   # :231-236
                    dst = free(src)
                    lib.ssr_warp_linear(src, dst, c['Tie'], None, B, *os_, 0, 0, 0, 0, 0, 0, None, *os_, None, st)
                    src = dst
                if src is not None:
                    lib.ssr_copy_strided(src, self.image, B * int(np.prod(os_)), 1, 0, p.n_image_channels, out_c, st)
            out_c += 1
            if p.build_reliability_maps:
                rel = c.get('rel')
                if c['reg']:                                                   # :237-238 warp the reliability map too
                    lib.ssr_fill_outer3(self.tmp_a, *(rel or [None, None, None]), B, *os_, 1, 0, st)
                    lib.ssr_warp_linear(self.tmp_a, self.tmp_b, c['Tie'], None, B, *os_, 0, 0, 0, 0, 0, 0, None, *os_,
                                        None, st)
                    lib.ssr_copy_strided(self.tmp_b, self.image, B * int(np.prod(os_)), 1, 0, p.n_image_channels,
                                         out_c, st)
                else:
                    lib.ssr_fill_outer3(self.image, *(rel or [None, None, None]), B, *os_, p.n_image_channels, out_c, st)
                out_c += 1
        cs = p.crop_shape
        if p.use_real_image:                                                   # :248-255
            lib.ssr_minmax(self.real, self.minmax, B, int(np.prod(cs)), st)
            unit = torch.ones(1, dtype=torch.float32, device=self.device)   # normalise through a 1x1x1 "stencil"
            lib.ssr_blur3d(self.real, self.tmp_a, unit, 1, 1, 1, self.minmax, None, B, *cs, 1, 0, 1, 0, st)
            self._emit_target(self.tmp_a, ktgt_real if p.target_sigma is not None else None, 0, st)
        return self.image, self.target

    def _emit_target(self, src, ktgt, tgt_c, st, rebound=False):
        """regression target: optional blur + linear resample to output_shape (labels_to_image_model.py:189-196)."""
        p, B = self.plan, self.B
        cs, os_ = p.crop_shape, p.output_shape
        nt = p.n_target_channels
        if os_ == cs:
            lib.ssr_copy_strided(src, self.target, B * int(np.prod(cs)), 1, 0, nt, tgt_c, st)
            return src
        cur = src
        bufs = [self.tmp_b, self.tmp_c]
        for n, (kp, ksh) in enumerate(ktgt or []):
            dst = bufs[n % 2]
            lib.ssr_blur3d(cur, dst, kp, *ksh, None, None, B, *cs, 1, 0, 1, 0, st)
            cur = dst
        if not rebound:
            lib.ssr_resize(cur, self.target, B, *cs, *os_, 1, 0, nt, tgt_c, st)
            return None
        # the input chain of this channel continues from the resampled target (GeneratorPlan.chan_grid): resample into a
        # contiguous [B, *output_shape] buffer first, then interleave it into the target tensor
        res = next(t for t in (self.tmp_b, self.tmp_c, self.tmp_a) if t is not cur and t is not src)
        lib.ssr_resize(cur, res, B, *cs, *os_, 1, 0, 0, 0, st)
        lib.ssr_copy_strided(res, self.target, B * int(np.prod(os_)), 1, 0, nt, tgt_c, st)
        return res
