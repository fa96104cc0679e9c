"""Host-side sampling of every random quantity the generator graph draws per step.

The reference draws these inside the TF graph (tf.random.*): ext/lab2im/utils.py:675-752,1019-1035 (affine),
ext/lab2im/layers.py:189-190 (SVF), :267 (crop), :400 (flip), :1080,1090 (bias), :1240 (gamma),
ext/lab2im/edit_tensors.py:119-121 (blur jitter), SynthSR/labels_to_image_model.py:205,233 (registration error).
They are tiny (a few hundred floats) except the per-voxel GMM noise, which is generated on the device (Philox) unless
`gmm_noise=True` asks for an injected volume (parity tests).  The resulting dict is the `draws` interface shared with
the oracle.
"""
import math

import numpy as np

f32 = np.float32


def _bounds(hyper, size, centre, default_range, rng=None):
    """-> (lo, hi) arrays of length `size` following draw_value_from_distribution (ext/lab2im/utils.py:1001-1016)."""
    if isinstance(hyper, str):
        hyper = np.load(hyper)
    if hyper is None:
        return np.full(size, centre - default_range), np.full(size, centre + default_range)
    if isinstance(hyper, np.ndarray):
        assert hyper.shape[0] % 2 == 0
        n_mod = hyper.shape[0] // 2
        idx = 2 * int((rng or np.random.default_rng()).integers(n_mod)) if n_mod > 1 else 0
        return hyper[idx], hyper[idx + 1]
    if isinstance(hyper, (int, float, np.integer, np.floating)):
        return np.full(size, centre - hyper), np.full(size, centre + hyper)
    if isinstance(hyper, (list, tuple)):
        assert len(hyper) == 2
        return np.full(size, hyper[0]), np.full(size, hyper[1])
    raise ValueError('bounds should be None, a number, a sequence of 2, or an array')


def _uniform(rng, hyper, batch, size, centre=0., default_range=10.):
    if hyper is False:
        return None
    lo, hi = _bounds(hyper, size, centre, default_range, rng)
    return rng.uniform(np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64), size=(batch, size)).astype(f32)


def sample_draws(rng, plan, batch, gmm_noise=False):
    """plan: synthsr_b200.generator.GeneratorPlan."""
    d = {}
    d['aff_rotation'] = _uniform(rng, plan.rotation_bounds, batch, 3, 0., 15.)
    d['aff_shearing'] = _uniform(rng, plan.shearing_bounds, batch, 6, 0., .01)
    d['aff_scaling'] = _uniform(rng, plan.scaling_bounds, batch, 3, 1., .15)
    d['aff_translation'] = _uniform(rng, plan.translation_bounds, batch, 3, 0., 5.)
    if plan.nonlin_std > 0:
        d['svf_std'] = f32(rng.uniform(0., plan.nonlin_std))
        d['svf_normal'] = rng.standard_normal((batch, *plan.svf_small_shape, 3), dtype=f32)
    if plan.crop_shape != plan.grid_shape:
        mx = np.array(plan.grid_shape) - np.array(plan.crop_shape)
        d['crop_idx'] = np.stack([(rng.uniform(0., mx)).astype(f32).astype(np.int32) for _ in range(batch)])
    else:
        d['crop_idx'] = np.zeros((batch, 3), np.int32)
    d['flip'] = (rng.uniform(0., 1., size=batch) < 0.5) if plan.flipping else np.zeros(batch, bool)
    if gmm_noise:
        d['gmm_normal'] = rng.standard_normal((batch, *plan.crop_shape, plan.n_channels), dtype=f32)
    for i in range(plan.n_channels):
        if plan.input_channels[i] and plan.bias_field_std > 0:
            d['bias_std_%d' % i] = rng.uniform(0., plan.bias_field_std, size=batch).astype(f32)
            d['bias_normal_%d' % i] = rng.standard_normal((batch, *plan.bias_small_shape), dtype=f32)
            d['bias_apply_%d' % i] = bool(rng.uniform() < 0.95)
        d['gamma_normal_%d' % i] = rng.standard_normal(batch, dtype=f32)
        if plan.input_channels[i]:
            r = plan.blur_range
            if r is not None and r != 1 and not plan.randomise_res[i]:    # GaussianBlur(sigma, blur_range) of the fixed-res branch
                d['blur_mult_%d' % i] = rng.uniform(1. / r, r, size=3).astype(f32)
            if plan.randomise_res[i]:
                # SampleResolution(atlas_res, max_res_iso=[9,9,9]) (ext/lab2im/layers.py:619-625, 646-647): resolution
                # U(atlas_res, 9) independently per (example, axis); with probability prob_min = 0.05 (one draw for the
                # whole batch) the atlas resolution; thickness U(atlas_res, resolution)
                lo = np.asarray(plan.atlas_res, dtype=f32)
                res = (lo + rng.uniform(0., 1., size=(batch, 3)).astype(f32) * (f32(9.) - lo)).astype(f32)
                if rng.uniform() < 0.05:
                    res = np.tile(lo[None], (batch, 1)).astype(f32)
                d['res_%d' % i] = res
                d['thick_%d' % i] = (lo + rng.uniform(0., 1., size=(batch, 3)).astype(f32) * (res - lo)).astype(f32)
                if r is not None and r != 1:   # gaussian_kernel jitter on the [B,3] sigma tensor (edit_tensors.py:119-121)
                    d['blur_mult_dyn_%d' % i] = rng.uniform(1. / r, r, size=(batch, 3)).astype(f32)
            if plan.sim_reg[i] and i != plan.idx_first_input_channel:
                d['reg_rot_%d' % i] = rng.uniform(-5., 5., size=(batch, 3)).astype(f32)
                d['reg_trans_%d' % i] = rng.uniform(-5., 5., size=(batch, 3)).astype(f32)
                d['reg_err_rot_%d' % i] = rng.uniform(-.5, .5, size=(batch, 3)).astype(f32)
                d['reg_err_trans_%d' % i] = rng.uniform(-.5, .5, size=(batch, 3)).astype(f32)
    return d


# ---------------------------------------------------------------------------------------------------------------------
# 4x4 affine assembly in float32 with a pinned, left-to-right accumulation order (the label output is bit exact only
# if the matrix is): T = [S . (Sh . R) | t] with R = Rx.Ry.Rz   (ext/lab2im/utils.py:735, 755-815)
# ---------------------------------------------------------------------------------------------------------------------
def _mm(a, b):
    n, m, p = a.shape[0], a.shape[1], b.shape[1]
    out = np.zeros((n, p), dtype=f32)
    for i in range(n):
        for j in range(p):
            acc = f32(a[i, 0] * b[0, j])
            for k in range(1, m):
                acc = f32(acc + f32(a[i, k] * b[k, j]))
            out[i, j] = acc
    return out


def rotation_matrix(rot_deg):
    r = (np.asarray(rot_deg, dtype=f32) * f32(np.pi)).astype(f32)
    r = (r / f32(180)).astype(f32)
    c, s = np.cos(r).astype(f32), np.sin(r).astype(f32)
    rx = np.array([[1, 0, 0], [0, c[0], -s[0]], [0, s[0], c[0]]], dtype=f32)
    ry = np.array([[c[1], 0, s[1]], [0, 1, 0], [-s[1], 0, c[1]]], dtype=f32)
    rz = np.array([[c[2], -s[2], 0], [s[2], c[2], 0], [0, 0, 1]], dtype=f32)
    return _mm(_mm(rx, ry), rz)


def build_affine(rotation=None, shearing=None, scaling=None, translation=None):
    rot = rotation_matrix(rotation) if rotation is not None else np.eye(3, dtype=f32)
    if shearing is not None:
        sh = np.asarray(shearing, dtype=f32)
        shm = np.array([[1, sh[0], sh[1]], [sh[2], 1, sh[3]], [sh[4], sh[5], 1]], dtype=f32)
    else:
        shm = np.eye(3, dtype=f32)
    sc = np.diag(np.asarray(scaling, dtype=f32)) if scaling is not None else np.eye(3, dtype=f32)
    t = np.eye(4, dtype=f32)
    t[:3, :3] = _mm(sc, _mm(shm, rot))
    if translation is not None:
        t[:3, 3] = np.asarray(translation, dtype=f32)
    return t


def matmul4(a, b):
    return _mm(np.asarray(a, dtype=f32), np.asarray(b, dtype=f32))


def resample_shape(shape, factor):
    factor = [factor] * len(shape) if np.isscalar(factor) else list(factor)
    return [math.ceil(shape[i] * factor[i]) for i in range(len(shape))]
