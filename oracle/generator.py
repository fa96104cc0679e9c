"""NumPy restatement of the reference's on-the-fly synthetic-scan generator (``labels_to_image_model``).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- PARITY UNPINNED (no reference tests / TF not installable).

Every function cites the reference file:line (relative to /root/reference) it restates.  All arithmetic is
float32 with the operation order spelled out, one IEEE rounding per operation (no fused multiply-add), because
the label-resampling output has to be compared bit-exactly with the CUDA kernels.  All randomness is
externalised in a ``draws`` dict (see ``synthsr_b200.draws`` for the producer) -- TensorFlow's Philox streams
are not reproducible without TensorFlow, so "identical seeds" is realised as "identical injected draws".

Layout: volumes are [X, Y, Z, C] (channels last, C-order), batches [B, X, Y, Z, C], like the reference.
"""
import itertools
import math

import numpy as np

f32 = np.float32


# ----------------------------------------------------------------------------------------------------------------
# ext/neuron/utils.py
# ----------------------------------------------------------------------------------------------------------------

def _grid(shape):
    """float32 ij-meshgrid, list of 3 arrays of shape `shape` (ext/neuron/utils.py:411-430, 449-521)."""
    return [g.astype(f32) for g in np.meshgrid(*[np.arange(s) for s in shape], indexing='ij')]


def interpn_linear(vol, loc):
    """ext/neuron/utils.py:67-110.  vol [X,Y,Z,C] f32, loc list of 3 f32 arrays (same shape S) -> S + [C]."""
    vol = np.asarray(vol, dtype=f32)
    shp = vol.shape[:3]
    mx = [f32(d - 1) for d in shp]
    loc = [np.asarray(l, dtype=f32) for l in loc]
    loc0 = [np.floor(l) for l in loc]                                           # :68
    clipped = [np.clip(loc[d], f32(0), mx[d]) for d in range(3)]                # :72
    loc0l = [np.clip(loc0[d], f32(0), mx[d]) for d in range(3)]                 # :73
    loc1 = [np.clip(loc0l[d] + f32(1), f32(0), mx[d]) for d in range(3)]        # :76
    locs = [[a.astype(np.int32) for a in loc0l], [a.astype(np.int32) for a in loc1]]
    diff1 = [(loc1[d] - clipped[d]).astype(f32) for d in range(3)]              # :82
    diff0 = [(f32(1) - diff1[d]).astype(f32) for d in range(3)]                 # :83
    wloc = [diff1, diff0]                                                       # :84
    flat = vol.reshape(-1, vol.shape[-1])
    out = None
    for c in itertools.product([0, 1], repeat=3):                               # :88-110
        idx = (locs[c[0]][0].astype(np.int64) * shp[1] + locs[c[1]][1]) * shp[2] + locs[c[2]][2]   # sub2ind :537-548
        val = flat[idx]
        wt = ((wloc[c[0]][0] * wloc[c[1]][1]).astype(f32) * wloc[c[2]][2]).astype(f32)   # prod_n :530-534
        term = (wt[..., None] * val).astype(f32)
        out = term if out is None else (out + term).astype(f32)                 # 0 + x == x exactly
    return out


def interpn_nearest(vol, loc):
    """ext/neuron/utils.py:112-122: tf.round (half-to-even) then clip, gather."""
    shp = vol.shape[:3]
    r = [np.rint(np.asarray(l, dtype=f32)).astype(np.int32) for l in loc]       # :114 tf.round = half to even
    r = [np.clip(r[d], 0, shp[d] - 1) for d in range(3)]                        # :117-118
    idx = (r[0].astype(np.int64) * shp[1] + r[1]) * shp[2] + r[2]
    return vol.reshape(-1, vol.shape[-1])[idx]


def transform(vol, shift, method='linear'):
    """ext/neuron/utils.py:289-320: sample vol at mesh + shift."""
    mesh = _grid(shift.shape[:3])
    loc = [(mesh[d] + shift[..., d].astype(f32)).astype(f32) for d in range(3)]  # :317
    return interpn_linear(vol, loc) if method == 'linear' else interpn_nearest(vol, loc)


def resize(vol, new_shape, method='linear'):
    """ext/neuron/layers.py:361-394 + ext/neuron/utils.py:127-154.

    zoom = size / inshape (python float, layers.py:379) is converted to a float32 constant when it divides the
    float32 grid; offset = grid / zoom - grid (utils.py:150); loc = grid + offset (utils.py:317)."""
    in_shape = vol.shape[:3]
    zoom = [f32(new_shape[d] / in_shape[d]) for d in range(3)]
    grid = _grid(new_shape)
    offset = [((grid[d] / zoom[d]).astype(f32) - grid[d]).astype(f32) for d in range(3)]
    return transform(vol, np.stack(offset, -1), method)


def integrate_vec(vec, nb_steps=7):
    """ext/neuron/utils.py:351-369 scaling and squaring (VecInt defaults: ext/neuron/layers.py:196)."""
    vec = (np.asarray(vec, dtype=f32) / f32(2 ** nb_steps)).astype(f32)
    for _ in range(nb_steps):
        vec = (vec + transform(vec, vec)).astype(f32)
    return vec


def affine_to_shift(aff, volshape, field=None):
    """ext/neuron/utils.py:160-219 (affine only) and :222-286 (non-linear field + affine).

    The 4x4 @ 4xV tf.matmul is pinned to left-to-right float32 accumulation without FMA:
    q = ((A0*p0 + A1*p1) + A2*p2) + A3*1."""
    aff = np.asarray(aff, dtype=f32)
    mesh = _grid(volshape)
    mesh_c = [(mesh[d] - f32((volshape[d] - 1) / 2)).astype(f32) for d in range(3)]          # :206 / :271
    if field is not None:
        p = [(mesh_c[d] + field[..., d].astype(f32)).astype(f32) for d in range(3)]          # :276
    else:
        p = mesh_c
    out = []
    for d in range(3):
        q = (aff[d, 0] * p[0]).astype(f32)
        q = (q + (aff[d, 1] * p[1]).astype(f32)).astype(f32)
        q = (q + (aff[d, 2] * p[2]).astype(f32)).astype(f32)
        q = (q + aff[d, 3]).astype(f32)                                                      # ones row * A[d,3]
        out.append((q - mesh_c[d]).astype(f32))                                              # :219 / :286
    return np.stack(out, -1)


def spatial_transformer(vol, aff=None, field=None, method='linear'):
    """ext/neuron/layers.py:125-179: non-linear first, affine second; pull-warp."""
    shape = vol.shape[:3]
    if aff is not None:
        shift = affine_to_shift(aff, shape, field)
    else:
        shift = field
    return transform(vol, shift, method)


# ----------------------------------------------------------------------------------------------------------------
# ext/lab2im/utils.py
# ----------------------------------------------------------------------------------------------------------------

def rotation_matrix(rot_deg):
    """ext/lab2im/utils.py:755-782: R = Rx @ Ry @ Rz, float32."""
    r = (np.asarray(rot_deg, dtype=f32) * f32(np.pi)).astype(f32)
    r = (r / f32(180)).astype(f32)                                                            # :757 rotation*pi/180
    c, s = np.cos(r).astype(f32), np.sin(r).astype(f32)
    Rx = np.array([[1, 0, 0], [0, c[0], -s[0]], [0, s[0], c[0]]], dtype=f32)
    Ry = np.array([[c[1], 0, s[1]], [0, 1, 0], [-s[1], 0, c[1]]], dtype=f32)
    Rz = np.array([[c[2], -s[2], 0], [s[2], c[2], 0], [0, 0, 1]], dtype=f32)
    return _mm3(_mm3(Rx, Ry), Rz)


def _mm3(a, b):
    """3x3 float32 matmul, left-to-right accumulation."""
    out = np.zeros((a.shape[0], b.shape[1]), dtype=f32)
    for i in range(a.shape[0]):
        for j in range(b.shape[1]):
            acc = f32(a[i, 0] * b[0, j])
            for k in range(1, a.shape[1]):
                acc = f32(acc + f32(a[i, k] * b[k, j]))
            out[i, j] = acc
    return out


def shearing_matrix(sh):
    """ext/lab2im/utils.py:797-807."""
    sh = np.asarray(sh, dtype=f32)
    return np.array([[1, sh[0], sh[1]], [sh[2], 1, sh[3]], [sh[4], sh[5], 1]], dtype=f32)


def build_affine(rotation=None, shearing=None, scaling=None, translation=None):
    """ext/lab2im/utils.py:675-752: T = [S @ (Sh @ R) | t ; 0 0 0 1]  (one example)."""
    R = rotation_matrix(rotation) if rotation is not None else np.eye(3, dtype=f32)
    Sh = shearing_matrix(shearing) if shearing is not None else np.eye(3, dtype=f32)
    S = np.diag(np.asarray(scaling, dtype=f32)) if scaling is not None else np.eye(3, dtype=f32)
    T3 = _mm3(S, _mm3(Sh, R))                                                                 # :735
    T = np.eye(4, dtype=f32)
    T[:3, :3] = T3
    if translation is not None:
        T[:3, 3] = np.asarray(translation, dtype=f32)
    return T


def get_resample_shape(shape, factor):
    """ext/lab2im/utils.py:577-588."""
    factor = _to_list(factor, len(shape))
    return [math.ceil(shape[i] * factor[i]) for i in range(len(shape))]


def _to_list(v, n):
    if isinstance(v, np.ndarray):
        v = np.squeeze(v).tolist()
    if isinstance(v, (int, float, np.integer, np.floating)):
        v = [v]
    v = list(v)
    if len(v) == 1:
        v = v * n
    assert len(v) == n
    return v


def find_closest_number_divisible_by_m(n, m):
    """ext/lab2im/utils.py:928-944, answer_type='lower'."""
    return n if n % m == 0 else int(n / m) * m


def get_mapping_lut(source, dest):
    """ext/lab2im/utils.py:894-914."""
    lut = np.zeros(int(np.max(source)) + 1, dtype=np.int32)
    for s, d in zip(source, dest):
        lut[s] = d
    return lut


# ----------------------------------------------------------------------------------------------------------------
# ext/lab2im/edit_tensors.py
# ----------------------------------------------------------------------------------------------------------------

def blurring_sigma_for_downsampling(current_res, downsample_res, mult_coef=None, thickness=None):
    """ext/lab2im/edit_tensors.py:41-83 (numpy branch)."""
    current_res = np.array(current_res, dtype=np.float64)
    downsample_res = np.array(downsample_res, dtype=np.float64)
    if thickness is not None:
        downsample_res = np.minimum(downsample_res, np.array(thickness, dtype=np.float64))
    if mult_coef is None:
        sigma = 0.75 * downsample_res / current_res
        sigma[downsample_res == current_res] = 0.5
    else:
        sigma = mult_coef * downsample_res / current_res
    sigma[downsample_res == 0] = 0
    return sigma


def gaussian_kernel_dense(sigma, blur_mult=None):
    """ext/lab2im/edit_tensors.py:86-181, separable=False branch, fixed sigma (list of 3).

    Window from the un-jittered sigma (:116,:124); jitter sigma*U(1/r,r) per axis (:119-121) given as blur_mult."""
    max_sigma = np.array(_to_list(sigma, 3), dtype=np.float64)
    sig = np.array(_to_list(sigma, 3), dtype=f32)
    if blur_mult is not None:
        sig = (sig * np.asarray(blur_mult, dtype=f32)).astype(f32)
    ws = np.int32(np.ceil(2.5 * max_sigma) / 2) * 2 + 1                                       # :124
    mesh = [g.astype(f32) for g in np.meshgrid(*[np.arange(w) for w in ws], indexing='ij')]
    diff = np.stack([(mesh[d] - f32((ws[d] - 1) / 2)).astype(f32) for d in range(3)], -1)    # :160
    is0 = sig == 0
    s1 = np.where(is0, f32(1), sig).astype(f32)
    exp_term = (-np.square(diff) / (f32(2) * s1 ** 2).astype(f32)).astype(f32)              # :174
    logt = np.log(np.where(is0, f32(1), (f32(np.sqrt(2 * np.pi)) * sig).astype(f32))).astype(f32)
    norms = (exp_term - logt).astype(f32)                                                    # :175
    k = np.exp(np.sum(norms, -1, dtype=f32)).astype(f32)                                     # :176-177
    k = (k / np.sum(k, dtype=f32)).astype(f32)                                               # :178
    return k


def gaussian_kernels_separable(sigma, blur_mult=None):
    """ext/lab2im/edit_tensors.py:126-154: list of three 1-D kernels (None when window is 1)."""
    max_sigma = np.array(_to_list(sigma, 3), dtype=np.float64)
    sig = np.array(_to_list(sigma, 3), dtype=f32)
    if blur_mult is not None:
        sig = (sig * np.asarray(blur_mult, dtype=f32)).astype(f32)
    ws = np.int32(np.ceil(2.5 * max_sigma) / 2) * 2 + 1
    out = []
    for i, w in enumerate(ws):
        if w > 1:
            loc = (np.arange(w).astype(f32) - f32((w - 1) / 2)).astype(f32)                  # :137
            exp_term = (-np.square(loc) / (f32(2) * sig[i] ** 2)).astype(f32)                # :145
            g = np.exp(exp_term - np.log(f32(np.sqrt(2 * np.pi)) * sig[i]).astype(f32)).astype(f32)
            g = (g / np.sum(g, dtype=f32)).astype(f32)                                        # :147
            out.append(g)
        else:
            out.append(None)
    return out


def conv3d_same(img, k):
    """tf.nn.conv3d(..., 'SAME') of a single-channel volume [X,Y,Z] with dense kernel k: correlation with zero
    padding (ext/lab2im/layers.py:748,758).  Accumulated in float64 then rounded (TF's own summation order is
    unknowable; compared with a tolerance)."""
    kx, ky, kz = k.shape
    px, py, pz = kx // 2, ky // 2, kz // 2
    pad = np.pad(img.astype(np.float64), ((px, px), (py, py), (pz, pz)))
    out = np.zeros(img.shape, dtype=np.float64)
    X, Y, Z = img.shape
    for a in range(kx):
        for b in range(ky):
            for c in range(kz):
                out += float(k[a, b, c]) * pad[a:a + X, b:b + Y, c:c + Z]
    return out.astype(f32)


def gaussian_blur(img, sigma, blur_mult=None):
    """ext/lab2im/layers.py:706-767 for one channel [X,Y,Z]: dense kernel if ||sigma|| <= 5 (:720) else three
    separable passes; skipped when all sigma are 0 (:757)."""
    sigma = _to_list(sigma, 3)
    separable = np.linalg.norm(np.array(sigma)) > 5
    if separable:
        ks = gaussian_kernels_separable(sigma, blur_mult)
        for ax, g in enumerate(ks):
            if g is not None:
                shape = [1, 1, 1]
                shape[ax] = len(g)
                img = conv3d_same(img, g.reshape(shape))
        return img
    if any(sigma):
        return conv3d_same(img, gaussian_kernel_dense(sigma, blur_mult))
    return img


def dynamic_sigma(atlas_res, resolution, thickness, mult_coef=.42):
    """ext/lab2im/edit_tensors.py:66-81 (tensor branch): sigma = mult * min(res, thickness) / current_res, 0 where the
    down-sampling resolution is 0.  resolution / thickness [B,3] float32."""
    res = np.asarray(resolution, dtype=f32)
    down = np.minimum(res, np.asarray(thickness, dtype=f32)).astype(f32)
    sigma = ((f32(mult_coef) * down).astype(f32) / np.asarray(atlas_res, dtype=f32)).astype(f32)
    return np.where(down == 0, f32(0), sigma).astype(f32)


def dynamic_separable_kernels(sigma, max_sigma, blur_mult=None):
    """ext/lab2im/edit_tensors.py:86-154 with sigma given as a [B,3] tensor (DynamicGaussianBlur, layers.py:813):
    window from max_sigma (:124), jitter sigma*U(1/r,r) per (example, axis) (:119-121), 1-D Gaussians per example, and --
    as the reference does (:147, `g / tf.reduce_sum(g)` on the [B, window] tensor) -- normalised by the sum over the
    WHOLE batch.  Returns three arrays [B, window]."""
    sig = np.asarray(sigma, dtype=f32)
    if blur_mult is not None:
        sig = (sig * np.asarray(blur_mult, dtype=f32)).astype(f32)
    max_sigma = np.array(_to_list(max_sigma, 3), dtype=np.float64)
    ws = np.int32(np.ceil(2.5 * max_sigma) / 2) * 2 + 1
    out = []
    for i, w in enumerate(ws):
        if w > 1:
            loc = (np.arange(w).astype(f32) - f32((w - 1) / 2)).astype(f32)[None, :]
            si = sig[:, i:i + 1]
            exp_term = (-np.square(loc) / (f32(2) * si ** 2).astype(f32)).astype(f32)
            g = np.exp(exp_term - np.log((f32(np.sqrt(2 * np.pi)) * si).astype(f32)).astype(f32)).astype(f32)
            g = (g / np.sum(g, dtype=f32)).astype(f32)
            out.append(g)
        else:
            out.append(None)
    return out


def mimic_acquisition_zooms(inshape, volume_res, subsample_res, resample_shape):
    """ext/lab2im/layers.py:935-939: acquisition grid shape and the two zoom factors for one example."""
    full = (np.array(inshape) * np.array(volume_res, dtype=np.float64)).astype(f32)
    down_shape = (full / np.asarray(subsample_res, dtype=f32)).astype(np.int32)                    # :935-936
    down_zoom = (down_shape / np.array(inshape)).astype(f32)                                      # :937
    up_zoom = (np.array(resample_shape, dtype=np.int32) / down_shape).astype(f32)                # :938
    return down_shape, down_zoom, up_zoom


def mimic_acquisition(vol, subsample_res, volume_res, min_subsample_res, resample_shape):
    """ext/lab2im/layers.py:921-987 for one example, build_dist_map=True, noise_std=0.  vol [X,Y,Z,1] -> (vol', dist)."""
    inshape = list(vol.shape[:3])
    res = np.asarray(subsample_res, dtype=f32)
    down_tensor_shape = np.int32(np.array(inshape) * np.array(volume_res) / np.array(min_subsample_res))   # :906
    _, down_zoom, up_zoom = mimic_acquisition_zooms(inshape, volume_res, res, resample_shape)
    grid = _grid(list(down_tensor_shape))
    down_loc = [np.clip((grid[d] / down_zoom[d]).astype(f32), f32(0), f32(inshape[d])) for d in range(3)]   # :942-946
    low = interpn_nearest(vol, down_loc)                                                                   # :947
    ugrid = _grid(list(resample_shape))
    up_loc = [(ugrid[d] / up_zoom[d]).astype(f32) for d in range(3)]                                       # :961-962
    out = interpn_linear(low, up_loc)                                                                      # :963
    dsq = None
    for d in range(3):                                                                                     # :973-986
        f_dist = (up_loc[d] - np.floor(up_loc[d])).astype(f32)
        c_dist = (np.ceil(up_loc[d]) - up_loc[d]).astype(f32)
        dd = (np.minimum(f_dist, c_dist) * res[d]).astype(f32)
        sq = np.square(dd).astype(f32)
        dsq = sq if dsq is None else (dsq + sq).astype(f32)
    return out, np.sqrt(dsq).astype(f32)[..., None]


def reliability_map(resample_shape, downsample_shape):
    """ext/lab2im/edit_tensors.py:313-329 -> float32 [X,Y,Z]."""
    up = np.array(resample_shape) / np.array(downsample_shape)
    rel = 1
    for i in range(3):
        loc_float = np.arange(0, resample_shape[i], up[i])
        loc_floor = np.int32(np.floor(loc_float))
        loc_ceil = np.int32(np.clip(loc_floor + 1, 0, resample_shape[i] - 1))
        tmp = np.zeros(resample_shape[i])
        tmp[loc_floor] = 1 - (loc_float - loc_floor)
        tmp[loc_ceil] = tmp[loc_ceil] + (loc_float - loc_floor)
        shape = [1, 1, 1]
        shape[i] = resample_shape[i]
        rel = rel * np.reshape(tmp, shape)
    return np.asarray(rel, dtype=f32)


def resample_tensor(img, resample_shape, subsample_res=None, volume_res=None, build_reliability_map=False):
    """ext/lab2im/edit_tensors.py:257-338 for one channel volume [X,Y,Z,1]."""
    tensor_shape = list(img.shape[:3])
    downsample_shape = tensor_shape
    resample_shape = list(resample_shape)
    if subsample_res is not None:
        subsample_res = [float(v) for v in subsample_res]
        volume_res = [float(v) for v in volume_res]
        if subsample_res != volume_res:
            downsample_shape = [int(tensor_shape[i] * volume_res[i] / subsample_res[i]) for i in range(3)]   # :295
            img = resize(img, downsample_shape, 'nearest')                                                  # :299
    if resample_shape != downsample_shape:
        img = resize(img, resample_shape, 'linear')                                                         # :304
    if build_reliability_map:
        if downsample_shape != tensor_shape:
            rel = reliability_map(resample_shape, downsample_shape)[..., None]
        else:
            rel = np.ones_like(img)                                                                         # :333
        return img, rel
    return img


# ----------------------------------------------------------------------------------------------------------------
# SynthSR/labels_to_image_model.py
# ----------------------------------------------------------------------------------------------------------------

def get_shapes(labels_shape, output_shape, atlas_res, target_res, padding_margin, output_div_by_n):
    """SynthSR/labels_to_image_model.py:269-335."""
    atlas_res = [float(v) for v in atlas_res]
    target_res = [float(v) for v in target_res]
    n_dims = 3
    labels_shape = list(labels_shape)
    if padding_margin is not None:
        padding_margin = [int(v) for v in _to_list(padding_margin, n_dims)]
        labels_shape = [labels_shape[i] + 2 * padding_margin[i] for i in range(n_dims)]
    resample_factor = [atlas_res[i] / float(target_res[i]) for i in range(n_dims)] if atlas_res != target_res else None
    if output_shape is not None:
        output_shape = [int(v) for v in _to_list(output_shape, n_dims)]
        if resample_factor is not None:
            output_shape = [min(int(labels_shape[i] * resample_factor[i]), output_shape[i]) for i in range(n_dims)]
        else:
            output_shape = [min(labels_shape[i], output_shape[i]) for i in range(n_dims)]
        if output_div_by_n is not None:
            output_shape = [find_closest_number_divisible_by_m(s, output_div_by_n) for s in output_shape]
        if resample_factor is not None:
            cropping_shape = [int(np.around(output_shape[i] / resample_factor[i], 0)) for i in range(n_dims)]
        else:
            cropping_shape = output_shape
    else:
        if output_div_by_n is not None:
            if resample_factor is not None:
                output_shape = [int(labels_shape[i] * resample_factor[i]) for i in range(n_dims)]
                output_shape = [find_closest_number_divisible_by_m(s, output_div_by_n) for s in output_shape]
                cropping_shape = [int(np.around(output_shape[i] / resample_factor[i], 0)) for i in range(n_dims)]
            else:
                cropping_shape = [find_closest_number_divisible_by_m(s, output_div_by_n) for s in labels_shape]
                output_shape = cropping_shape
        else:
            cropping_shape = labels_shape
            if resample_factor is not None:
                output_shape = [int(cropping_shape[i] * resample_factor[i]) for i in range(n_dims)]
            else:
                output_shape = cropping_shape
    return cropping_shape, output_shape, padding_margin


def _n_channels_array(var, n_channels):
    """ext/lab2im/utils.py:373-397."""
    if var is None:
        return None
    var = np.array(var, dtype=np.float64)
    if var.ndim == 0:
        var = np.tile(var.reshape(1, 1), (n_channels, 3))
    elif var.ndim == 1:
        var = np.tile(var.reshape(1, 3), (n_channels, 1))
    return np.round(var.reshape(n_channels, 3), 3)


def resolve_config(cfg, labels_shape):
    """Static bookkeeping of SynthSR/labels_to_image_model.py:69-100 -> dict of derived quantities."""
    ic = cfg.get('input_channels', True)
    input_channels = [bool(ic)] if isinstance(ic, (bool, int, np.bool_)) else [bool(v) for v in np.asarray(ic).ravel()]
    n_channels = len(input_channels)
    output_channel = cfg.get('output_channel', 0)
    if output_channel is not None and not isinstance(output_channel, (list, tuple)):
        output_channel = [int(output_channel)]
    use_real_image = output_channel is None
    sim_reg = cfg.get('simulate_registration_error', True)
    sim_reg = [bool(sim_reg)] * n_channels if isinstance(sim_reg, (bool, int)) else [bool(v) for v in sim_reg]
    atlas_res = _n_channels_array(cfg.get('atlas_res', 1.), n_channels)
    data_res = cfg.get('data_res', None)
    thickness = cfg.get('thickness', None)
    if data_res is not None:
        data_res = np.array(data_res, dtype=np.float64).reshape(-1, 3) if np.ndim(data_res) > 0 else data_res
    if thickness is not None:
        thickness = np.array(thickness, dtype=np.float64).reshape(-1, 3) if np.ndim(thickness) > 0 else thickness
    if output_channel is not None:
        for idx in output_channel:
            if not input_channels[idx]:
                data_res = np.insert(data_res, idx, 1, axis=0)                                   # :88
                thickness = np.insert(thickness, idx, 1, axis=0)                                 # :89
    data_res = atlas_res if data_res is None else _n_channels_array(data_res, n_channels)      # :90
    thickness = data_res if thickness is None else _n_channels_array(thickness, n_channels)    # :91
    downsample = cfg.get('downsample', False)
    if downsample:
        downsample = [bool(downsample)] * n_channels if isinstance(downsample, (bool, int)) else list(downsample)
    else:
        downsample = list(np.min(thickness - data_res, 1) < 0)                                  # :92
    atlas_res0 = atlas_res[0]
    target_res = cfg.get('target_res', None)
    target_res = atlas_res0 if target_res is None else _n_channels_array(target_res, 1)[0]     # :94
    crop_shape, output_shape, padding_margin = get_shapes(labels_shape, cfg.get('output_shape', None), atlas_res0,
                                                          target_res, cfg.get('padding_margin', None),
                                                          cfg.get('output_div_by_n', None))
    return dict(input_channels=input_channels, n_channels=n_channels, output_channel=output_channel,
                use_real_image=use_real_image, sim_reg=sim_reg, atlas_res=atlas_res0, data_res=data_res,
                thickness=thickness, downsample=downsample, target_res=target_res, crop_shape=list(crop_shape),
                output_shape=list(output_shape), padding_margin=padding_margin,
                idx_first_input_channel=int(np.argmax(input_channels)))


def random_spatial_deformation_field(draws, b, inshape, nonlin_scale):
    """ext/lab2im/layers.py:185-197: SVF -> half-res resize -> VecInt -> full-res resize.  Returns (field, parts)."""
    small_shape = get_resample_shape(inshape, nonlin_scale)                                    # layers.py:151
    std = f32(draws['svf_std'])
    svf = (np.asarray(draws['svf_normal'][b], dtype=f32) * std).astype(f32)                    # :189-190
    assert list(svf.shape[:3]) == small_shape, (svf.shape, small_shape)
    resize_shape = [max(int(inshape[i] / 2), small_shape[i]) for i in range(3)]                # :193
    half = resize(svf, resize_shape, 'linear')                                                 # :194
    integ = integrate_vec(half, 7)                                                             # :195
    full = resize(integ, inshape, 'linear')                                                    # :196
    return full, dict(svf=svf, half=half, integrated=integ)


def labels_to_image(cfg, inputs, draws, return_intermediates=False):
    """One evaluation of the graph built by SynthSR/labels_to_image_model.py:32-266.

    cfg: dict of the reference's keyword arguments (generation_labels, n_neutral_labels, atlas_res, ...).
    inputs: [labels [B,X,Y,Z,1] int, means [B,L,C], stds [B,L,C] (, real_image [B,X,Y,Z,1])].
    Returns (image [B,*,Cimg] f32, target [B,*,Cout] f32) (+ dict of intermediates)."""
    labels_in = np.asarray(inputs[0])
    means_in = np.asarray(inputs[1], dtype=f32)
    stds_in = np.asarray(inputs[2], dtype=f32)
    B = labels_in.shape[0]
    labels_shape = list(labels_in.shape[1:4])
    r = resolve_config(cfg, labels_shape)
    gen_labels = np.asarray(cfg['generation_labels']).astype(np.int64)
    n_neutral = cfg.get('n_neutral_labels', None)
    n_neutral = len(gen_labels) if n_neutral is None else n_neutral
    rr = cfg.get('randomise_res', False)
    randomise_res = [bool(rr)] * r['n_channels'] if isinstance(rr, (bool, np.bool_)) or rr is None else [bool(v) for v in rr]
    inter = {}

    images, targets = [], []
    for b in range(B):
        lab = labels_in[b, ..., 0].astype(np.int32)
        real = np.asarray(inputs[3][b], dtype=f32) if r['use_real_image'] else None
        # --- pad (layers.py:1754-1755) ---
        if r['padding_margin'] is not None:
            pm = r['padding_margin']
            lab = np.pad(lab, [(m, m) for m in pm])
            if real is not None:
                real = np.pad(real, [(m, m) for m in pm] + [(0, 0)])
        shape = list(lab.shape)
        # --- RandomSpatialDeformation (layers.py:161-211) ---
        aff = None
        if any(cfg.get(k, d) is not False for k, d in [('scaling_bounds', .15), ('rotation_bounds', 15),
                                                       ('shearing_bounds', .012), ('translation_bounds', False)]):
            aff = build_affine(draws['aff_rotation'][b] if draws.get('aff_rotation') is not None else None,
                               draws['aff_shearing'][b] if draws.get('aff_shearing') is not None else None,
                               draws['aff_scaling'][b] if draws.get('aff_scaling') is not None else None,
                               draws['aff_translation'][b] if draws.get('aff_translation') is not None else None)
        field = None
        if cfg.get('nonlin_std', 3.) > 0:
            field, parts = random_spatial_deformation_field(draws, b, shape, cfg.get('nonlin_shape_factor', .0625))
            if b == 0:
                inter.update(parts)
                inter['field'] = field
        if aff is not None or field is not None:
            labf = lab.astype(f32)[..., None]                                                    # layers.py:167
            if aff is not None:
                labf = spatial_transformer(labf, aff, field, 'nearest')                          # :202
                if real is not None:
                    real = spatial_transformer(real, aff, field, 'linear')
            else:
                labf = transform(labf, field, 'nearest')
                if real is not None:
                    real = transform(real, field, 'linear')
            lab = labf[..., 0].astype(np.int32)                                                  # :209
        if b == 0:
            inter['affine'] = aff
            inter['labels_deformed'] = lab.copy()
        # --- RandomCrop (layers.py:252-270) ---
        cs = r['crop_shape']
        if cs != shape:
            lab = random_crop(lab, draws['crop_idx'][b], cs)
            if real is not None:
                real = random_crop(real, draws['crop_idx'][b], cs)
        # --- RandomFlip axis 0, swap labels (layers.py:362-427) ---
        if cfg.get('flipping', True):
            flip = bool(draws['flip'][b])
            lab = random_flip(lab, flip, gen_labels, n_neutral, swap=True)
            if real is not None:
                real = random_flip(real, flip, gen_labels, n_neutral, swap=False)
        lab = np.ascontiguousarray(lab)
        if b == 0:
            inter['labels'] = lab.copy()
        # --- SampleConditionalGMM (layers.py:472-498) ---
        C = r['n_channels']
        chans = [sample_conditional_gmm(lab, means_in[:, :, i], stds_in[:, :, i],
                                        np.asarray(draws['gmm_normal'][b, ..., i], dtype=f32), gen_labels) for i in range(C)]
        if b == 0:
            inter['gmm'] = np.stack(chans, -1)
        # --- per-channel processing (labels_to_image_model.py:175-242) ---
        out_channels, out_targets = [], []
        for i in range(C):
            ch = chans[i]
            if r['input_channels'][i]:
                ch = bias_field_corruption(ch, draws, b, i, cfg.get('bias_field_std', .3),
                                           cfg.get('bias_shape_factor', .025))                   # :180
                if b == 0:
                    inter['biased_%d' % i] = ch.copy()
            ch = intensity_augmentation(ch, clip=300, gamma=f32(f32(draws['gamma_normal_%d' % i][b]) * f32(.5)))  # :184
            if b == 0:
                inter['intensity_%d' % i] = ch.copy()
            ch = gaussian_blur(ch, .5)                                                           # :186
            if b == 0:
                inter['blur_%d' % i] = ch.copy()
            if not r['use_real_image'] and any(c == i for c in r['output_channel']):
                tgt = ch
                if r['crop_shape'] != r['output_shape']:
                    sigma = blurring_sigma_for_downsampling(r['atlas_res'], r['target_res'])     # :192
                    tgt = gaussian_blur(tgt, list(sigma))
                    tgt = resample_tensor(tgt[..., None], r['output_shape'])[..., 0]             # :195
                    # the reference REBINDS `channel` at :194-195: when this channel is also an input, its acquisition
                    # chain (:199-238) starts from the blurred, resampled target on the output grid, not from the
                    # full-resolution channel (pinned by executing the graph: tests/golden/make_reference_model_goldens.py)
                    ch = tgt
                out_targets.append(tgt)
            if r['input_channels'][i]:
                do_reg = r['sim_reg'][i] and (i != r['idx_first_input_channel'])
                Tinv = None
                if do_reg:                                                                       # :202-208
                    T = build_affine(rotation=draws['reg_rot_%d' % i][b], translation=draws['reg_trans_%d' % i][b])
                    Tinv = np.linalg.inv(T.astype(np.float64)).astype(f32)                       # tf.linalg.inv
                    ch = spatial_transformer(ch[..., None], T, None, 'linear')[..., 0]
                blur_range = cfg.get('blur_range', 1.15)
                if randomise_res[i]:                                                             # :215-220
                    max_res = np.array([9.] * 3)
                    res_all, thick_all = draws['res_%d' % i], draws['thick_%d' % i]              # SampleResolution
                    sig_all = dynamic_sigma(r['atlas_res'], res_all, thick_all, .42)
                    mult = draws['blur_mult_dyn_%d' % i] if (blur_range is not None and blur_range != 1) else None
                    ks = dynamic_separable_kernels(sig_all, 0.75 * max_res / np.array(r['atlas_res']), mult)
                    for ax, g in enumerate(ks):                                                  # layers.py:814-816
                        if g is not None:
                            kshape = [1, 1, 1]
                            kshape[ax] = g.shape[1]
                            ch = conv3d_same(ch, g[b].reshape(kshape))
                    ch, rel = mimic_acquisition(ch[..., None], res_all[b], r['atlas_res'], r['atlas_res'],
                                                r['output_shape'])
                    if b == 0:
                        inter['mimic_%d' % i] = ch.copy()
                else:
                    sigma = blurring_sigma_for_downsampling(r['atlas_res'], r['data_res'][i], .42, r['thickness'][i])  # :223
                    mult = draws['blur_mult_%d' % i] if (blur_range is not None and blur_range != 1) else None
                    ch = gaussian_blur(ch, list(sigma), mult)                                    # :224
                    if r['downsample'][i]:
                        ch, rel = resample_tensor(ch[..., None], r['output_shape'], list(r['data_res'][i]),
                                                  list(r['atlas_res']), True)                    # :226
                    else:
                        ch, rel = resample_tensor(ch[..., None], r['output_shape'], build_reliability_map=True)   # :228
                if do_reg:                                                                       # :231-238
                    Terr = build_affine(rotation=draws['reg_err_rot_%d' % i][b],
                                        translation=draws['reg_err_trans_%d' % i][b])
                    Tie = _mm4(Terr, Tinv)
                    ch = spatial_transformer(ch, Tie, None, 'linear')
                    rel = spatial_transformer(rel, Tie, None, 'linear')
                out_channels.append(ch[..., 0])
                if cfg.get('build_reliability_maps', False):
                    out_channels.append(rel[..., 0])
        image = np.stack(out_channels, -1).astype(f32)                                           # :245
        if r['use_real_image']:                                                                  # :248-255
            tgt = intensity_augmentation(real[..., 0], clip=0, gamma=None)
            if r['crop_shape'] != r['output_shape']:
                sigma = blurring_sigma_for_downsampling(r['atlas_res'], r['target_res'])
                tgt = gaussian_blur(tgt, list(sigma))
                tgt = resample_tensor(tgt[..., None], r['output_shape'])[..., 0]
            target = tgt[..., None].astype(f32)
        else:
            target = np.stack(out_targets, -1).astype(f32)                                       # :257
        images.append(image)
        targets.append(target)
    image = np.stack(images, 0)
    target = np.stack(targets, 0)
    if return_intermediates:
        return image, target, inter
    return image, target


def sample_conditional_gmm(lab, means, stds, noise, gen_labels):
    """ext/lab2im/layers.py:480-498 for one channel of one example: lab [X,Y,Z] int, means / stds [B, L] (the WHOLE batch:
    tf.scatter_nd adds the tables of all batch elements into one, which is then tiled over the batch -- with batchsize 1
    that is the per-example table), noise [X,Y,Z] standard normals.  Unlisted label values get mean = std = 0."""
    means, stds = np.asarray(means, dtype=f32), np.asarray(stds, dtype=f32)
    max_label = int(np.max(gen_labels)) + 1
    msum, ssum = means[0].copy(), stds[0].copy()
    for bb in range(1, means.shape[0]):
        msum = (msum + means[bb]).astype(f32)
        ssum = (ssum + stds[bb]).astype(f32)
    mlut = np.zeros(max_label, dtype=f32)
    slut = np.zeros(max_label, dtype=f32)
    mlut[gen_labels] = msum
    slut[gen_labels] = ssum
    return ((slut[lab] * noise).astype(f32) + mlut[lab]).astype(f32)                            # :498


def random_crop(vol, crop_idx, crop_shape):
    """ext/lab2im/layers.py:266-270: tf.slice at the (truncated) per-example offsets, same for every input."""
    ci = [int(v) for v in crop_idx]
    return vol[ci[0]:ci[0] + crop_shape[0], ci[1]:ci[1] + crop_shape[1], ci[2]:ci[2] + crop_shape[2]]


def random_flip(vol, flip, label_list, n_neutral, swap):
    """ext/lab2im/layers.py:362-427 for flip_axis 0: swap the right / left label values (inputs with swap_labels) when the
    example is flipped an odd number of times, then reverse axis 0."""
    if flip and swap and n_neutral != len(label_list):
        n_lab = len(label_list)
        split = np.split(np.asarray(label_list), [n_neutral, n_neutral + int((n_lab - n_neutral) / 2)])   # :382
        lut = get_mapping_lut(np.asarray(label_list), np.concatenate((split[0], split[2], split[1])))
        vol = lut[vol]
    return vol[::-1] if flip else vol


def _mm4(a, b):
    out = np.zeros((4, 4), dtype=f32)
    for i in range(4):
        for j in range(4):
            acc = f32(a[i, 0] * b[0, j])
            for k in range(1, 4):
                acc = f32(acc + f32(a[i, k] * b[k, j]))
            out[i, j] = acc
    return out


def bias_field_corruption(ch, draws, b, i, bias_field_std, bias_scale):
    """ext/lab2im/layers.py:1067-1097 for one channel [X,Y,Z] (prob=.95 -> 'bias_apply_i' draw)."""
    if not bias_field_std > 0:
        return ch
    small_shape = get_resample_shape(list(ch.shape), bias_scale)                                # :1059
    std = f32(draws['bias_std_%d' % i][b])
    small = (np.asarray(draws['bias_normal_%d' % i][b], dtype=f32) * std).astype(f32)          # :1080
    assert list(small.shape) == small_shape, (small.shape, small_shape)
    bias = resize(small[..., None], list(ch.shape), 'linear')[..., 0]                           # :1083
    bias = np.exp(bias).astype(f32)                                                             # :1084
    if bool(draws['bias_apply_%d' % i]):                                                        # :1090
        return (bias * ch).astype(f32)
    return ch


def intensity_augmentation(ch, clip, gamma):
    """ext/lab2im/layers.py:1186-1257 with noise_std=0, normalise=True, norm_perc=0, separate channels."""
    ch = np.asarray(ch, dtype=f32)
    if clip:
        ch = np.clip(ch, f32(0), f32(clip))                                                     # :1215
    m, M = ch.min(), ch.max()                                                                   # :1230-1231
    ch = np.clip(ch, m, M)
    ch = ((ch - m) / ((M - m).astype(f32) + f32(1e-7)).astype(f32)).astype(f32)                 # :1236
    if gamma is not None:
        ch = np.power(ch, np.exp(f32(gamma)).astype(f32)).astype(f32)                           # :1242
    return ch
