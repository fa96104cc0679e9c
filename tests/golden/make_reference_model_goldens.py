"""Golden vectors for the WHOLE generator graph: the reference's own SynthSR/labels_to_image_model.labels_to_image_model()
is executed, unmodified, on the NumPy `tf` shim -- KL.Input returns the fed array, every layer / Lambda runs eagerly in the
order the graph is built (which is the order TF evaluates the random ops' dependencies in), Model() just holds the outputs.
Every tf.random draw is generated here from a seeded NumPy generator, logged in call order, and stored next to the
outputs, so the test can hand the SAME numbers to oracle.generator.labels_to_image through its `draws` interface.

Eight configurations (all tiny, the shim is pure NumPy):
  A  training() defaults shape-for-shape: 1 channel, crop, flip, elastic + affine, bias, gamma, blur jitter,
     anisotropic acquisition (data_res [1,1,3], thickness [1,1,2] -> downsample), reliability map;
  B  batch 2, 2 channels: second channel with simulated registration error and randomise_res (SampleResolution,
     DynamicGaussianBlur, MimicAcquisition), first channel also the regression target at target_res 2 (blur + resample);
  C  real image as regression target (output_channel None), padding margin, no flipping, no elastic part;
  D  target-only second channel at target_res 2 (blur + resample), anisotropic input channel resampled to the output grid;
  E  second channel both input and target at target_res 1.5: the reference rebinds `channel` to the resampled target
     (:194-195), so that channel's registration error, acquisition blur, down/up-sampling and reliability map all run
     on the OUTPUT grid;
  F  batch 2 with per-example crops, rotation + translation only, no flipping, bias field drawn but skipped (the 5 %
     branch), no blur jitter,
     blur-only acquisition (data_res [2,1,1], no down-sampling -> all-ones reliability map);
  G  no spatial deformation at all, asymmetric padding margin + output_div_by_n, three channels (target-only + two inputs,
     the second with registration error), thickness above / below the slice spacing, blur_range 1, no reliability maps;
  H  randomise_res with the 5 % branch taken (acquisition at the atlas resolution).

Writes tests/golden/reference_model.npz.   (build container only: needs /root/reference)"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_numpy_shim_layers as shim  # noqa: E402

tf, K, T = shim.install([])
f32 = np.float32
_np = np.asarray

# ---- the few extra pieces the graph-building function needs on top of the layer shim ----------------------------------
FEED, LOG, FORCED = [], [], []
BOUNDS = {}
RNG = [None]


def _shape(s):
    return tuple(int(v) for v in _np(s).reshape(-1))


def uniform(shape, minval=0, maxval=1, dtype='float32'):
    shp = _shape(shape)
    if 'int' in str(dtype):
        v = RNG[0].integers(int(minval), int(maxval), size=shp).astype(np.int32)
        LOG.append(('uniform_int', v, minval, maxval))
        return T(v)
    lo, hi = _np(_np(minval), dtype=f32), _np(_np(maxval), dtype=f32)
    if shp == (1,) and float(lo) == 0. and float(hi) == 1. and FORCED:
        u = np.array([FORCED.pop(0)], dtype=f32)                      # branch probabilities chosen by the script
    else:
        u = RNG[0].uniform(size=shp).astype(f32)
    v = (u * (hi - lo).astype(f32) + lo).astype(f32)                  # the TF kernel: rand * (maxval - minval) + minval
    v = np.minimum(v, np.nextafter(np.broadcast_to(hi, v.shape), -np.inf, dtype=f32)) if np.any(v >= hi) and np.all(hi > lo) else v
    LOG.append(('uniform', v, lo.tolist(), hi.tolist()))
    return T(v)


def normal(shape, mean=0., stddev=1., dtype='float32'):
    z = RNG[0].standard_normal(_shape(shape)).astype(f32)
    LOG.append(('normal', z, _np(_np(mean), dtype=f32).tolist(), _np(_np(stddev), dtype=f32).tolist()))   # the STANDARD normal is what the oracle gets
    return T((z * _np(_np(stddev), dtype=f32) + _np(_np(mean), dtype=f32)).astype(f32))


def tensor_scatter_nd_update(tensor, indices, updates):
    out = _np(tensor).copy()
    idx = _np(indices)
    out[tuple(idx[..., d] for d in range(idx.shape[-1]))] = _np(updates)
    return T(out)


tf.random = types.SimpleNamespace(uniform=uniform, normal=normal)
tf.tensor_scatter_nd_update = tensor_scatter_nd_update
tf.linalg.inv = lambda x: T(np.linalg.inv(_np(x).astype(np.float64)).astype(f32))
tf.pad = lambda x, paddings, mode='CONSTANT', constant_values=0: T(np.pad(_np(x), [tuple(int(v) for v in r) for r in _np(paddings)],
                                                                        constant_values=constant_values))
tf.math.pow = lambda a, b: T(np.power(_np(a), _np(b)))
KLm = sys.modules['keras.layers']
KLm.Input = lambda shape=None, name=None, dtype=None: FEED.pop(0)


class Model:
    def __init__(self, inputs=None, outputs=None, name=None):
        self.inputs, self.outputs = inputs, outputs
        self.input = inputs[0] if isinstance(inputs, list) and len(inputs) == 1 else inputs
        self.output = outputs[0] if isinstance(outputs, list) and len(outputs) == 1 else outputs

    def get_layer(self, name):                               # (whole-training harness: layers are registered by name there)
        return sys.modules['make_reference_unet_goldens'].LAYERS[name]


sys.modules['keras.models'].Model = Model
sys.path.insert(0, '/root/reference')
from SynthSR.labels_to_image_model import labels_to_image_model  # noqa: E402

GEN = np.array([0, 14, 15, 16, 24, 2, 3, 4, 41, 42, 43])            # 5 neutral labels + 3 left/right pairs
N_NEUTRAL = 5


def run(seed, labels_shape, batch, forced, real=False, **kw):
    rng = np.random.default_rng(seed)
    RNG[0] = rng
    del LOG[:], FEED[:], FORCED[:]
    FORCED.extend(forced)
    C = len(kw['input_channels'])
    # blobby label map: a smooth random field quantised into the label list, so the deformed map has structure
    g = np.stack(np.meshgrid(*[np.linspace(0, 3, s) for s in labels_shape], indexing='ij'), -1)
    labs = []
    for b in range(batch):
        ph = rng.uniform(0, 6, size=3)
        fld = np.sin(g[..., 0] * 2 + ph[0]) + np.cos(g[..., 1] * 3 + ph[1]) + np.sin(g[..., 2] * 2.5 + ph[2])
        labs.append(GEN[np.clip(((fld + 3) / 6 * len(GEN)).astype(int), 0, len(GEN) - 1)])
    labels = np.stack(labs)[..., None].astype(np.int32)
    means = rng.uniform(20, 230, size=(batch, len(GEN), C)).astype(f32)
    stds = rng.uniform(1, 25, size=(batch, len(GEN), C)).astype(f32)
    inputs = [labels, means, stds]
    if real:
        inputs.append(rng.uniform(0, 200, size=labels.shape).astype(f32))
    FEED.extend(T(a) for a in inputs)
    shim.base.GRAPH_BATCH[0] = batch
    model = labels_to_image_model(labels_shape=list(labels_shape), generation_labels=GEN, n_neutral_labels=N_NEUTRAL, **kw)
    shim.base.GRAPH_BATCH[0] = None
    assert not FEED and not FORCED
    image, target = (np.asarray(o) for o in model.outputs)
    return inputs, image, target, [(k, np.asarray(v), a, b) for k, v, a, b in LOG]


def to_draws(cfg, log, batch, crop_differs):
    """call-order log -> the `draws` dict of synthsr_b200/draws.py (same order as the graph: labels_to_image_model.py:128-238)."""
    it = iter(log)
    d = {}
    last = [None]

    def nxt(kind):
        k, v, a, b = next(it)
        assert k == kind, (k, kind)
        last[0] = [k, a, b]                  # (minval, maxval) of a uniform, (mean, stddev) of a normal, as the graph passed them
        return v

    class Recorder(dict):                    # remembers the distribution parameters of the call that produced each entry
        def __setitem__(self, key, value):
            if value is not None and key not in BOUNDS:
                BOUNDS[key] = last[0]
            dict.__setitem__(self, key, value)

    BOUNDS.clear()
    d = Recorder()

    for key, name in [('rotation_bounds', 'aff_rotation'), ('shearing_bounds', 'aff_shearing'),
                      ('scaling_bounds', 'aff_scaling'), ('translation_bounds', 'aff_translation')]:
        d[name] = nxt('uniform') if cfg.get(key, {'translation_bounds': False}.get(key, True)) is not False else None
    if cfg.get('nonlin_std', 3.) > 0:
        d['svf_std'] = f32(nxt('uniform').reshape(-1)[0])
        d['svf_normal'] = nxt('normal')
    if crop_differs:
        d['crop_idx'] = np.stack([nxt('uniform').astype(np.int32) for _ in range(batch)])       # tf.cast -> int32 truncates
    else:
        d['crop_idx'] = np.zeros((batch, 3), np.int32)
        BOUNDS.pop('crop_idx', None)                                    # nothing was drawn
    if cfg.get('flipping', True):
        d['flip'] = nxt('uniform')[:, 0] < f32(.5)
    else:
        d['flip'] = np.zeros(batch, bool)
        BOUNDS.pop('flip', None)
    d['gmm_normal'] = nxt('normal')
    C = len(cfg['input_channels'])
    first = int(np.argmax(cfg['input_channels']))
    rr = cfg.get('randomise_res', False)
    rr = [rr] * C if isinstance(rr, bool) else rr
    sim = cfg.get('simulate_registration_error', True)
    sim = [sim] * C if isinstance(sim, bool) else sim
    jitter = cfg.get('blur_range', 1.15) is not None and cfg.get('blur_range', 1.15) != 1
    for i in range(C):
        inp = bool(cfg['input_channels'][i])
        if inp and cfg.get('bias_field_std', .3) > 0:
            d['bias_std_%d' % i] = nxt('uniform').reshape(batch)
            d['bias_normal_%d' % i] = nxt('normal')[..., 0]
            d['bias_apply_%d' % i] = bool(nxt('uniform')[0] < f32(.95))
        d['gamma_normal_%d' % i] = nxt('normal').reshape(batch)
        if inp:
            reg = sim[i] and i != first
            if reg:
                d['reg_rot_%d' % i] = nxt('uniform')
                d['reg_trans_%d' % i] = nxt('uniform')
            if rr[i]:
                nxt('uniform_int')                                     # anisotropy axis: drawn, unused when max_res_aniso is None
                res = nxt('uniform')
                BOUNDS['res_%d' % i] = last[0]
                at_min = bool(nxt('uniform')[0] < f32(.05))
                lo = np.tile(np.asarray(cfg['atlas_res'], dtype=f32).reshape(-1)[:3][None] if np.ndim(cfg['atlas_res']) else
                             np.full((1, 3), cfg['atlas_res'], f32), (batch, 1))
                d['res_%d' % i] = lo.astype(f32) if at_min else res
                d['thick_%d' % i] = nxt('uniform')
                if jitter:
                    d['blur_mult_dyn_%d' % i] = nxt('uniform')
            elif jitter:
                d['blur_mult_%d' % i] = nxt('uniform')
            if reg:
                d['reg_err_rot_%d' % i] = nxt('uniform')
                d['reg_err_trans_%d' % i] = nxt('uniform')
    rest = list(it)
    assert not rest, 'unconsumed draws: %s' % [(k, v.shape) for k, v, _, _ in rest]
    return dict(d)


CASES = {
    'A': dict(seed=101, labels_shape=(28, 32, 24), batch=1, forced=[.3], cfg=dict(
        input_channels=[True], output_channel=[0], atlas_res=1., target_res=1., output_shape=[24, 24, 16], flipping=True,
        aff=np.eye(4), scaling_bounds=.15, rotation_bounds=15, shearing_bounds=.012, translation_bounds=False, nonlin_std=3.,
        nonlin_shape_factor=.0625, simulate_registration_error=True, randomise_res=False, data_res=np.array([[1., 1., 3.]]),
        thickness=np.array([[1., 1., 2.]]), downsample=False, build_reliability_maps=True, blur_range=1.15, bias_field_std=.3,
        bias_shape_factor=.025)),
    'B': dict(seed=202, labels_shape=(24, 24, 32), batch=2, forced=[.5, .99, .6], cfg=dict(
        input_channels=[True, True], output_channel=[0], atlas_res=1., target_res=2., output_shape=None, flipping=True,
        aff=np.eye(4), scaling_bounds=.1, rotation_bounds=10, shearing_bounds=.01, translation_bounds=3, nonlin_std=2.,
        nonlin_shape_factor=.0625, simulate_registration_error=True, randomise_res=[False, True],
        data_res=np.array([[1., 1., 1.], [1., 1., 1.]]), thickness=None, downsample=False, build_reliability_maps=True,
        blur_range=1.15, bias_field_std=.3, bias_shape_factor=.025)),
    'C': dict(seed=303, labels_shape=(20, 16, 20), batch=1, forced=[.2], real=True, cfg=dict(
        input_channels=[True], output_channel=None, atlas_res=1., target_res=1., output_shape=None, padding_margin=2,
        flipping=False, aff=None, scaling_bounds=.15, rotation_bounds=15, shearing_bounds=.012, translation_bounds=False,
        nonlin_std=0., simulate_registration_error=False, randomise_res=False, data_res=np.array([[1., 4., 1.]]),
        thickness=np.array([[1., 4., 1.]]), downsample=True, build_reliability_maps=False, blur_range=1.15,
        bias_field_std=.3, bias_shape_factor=.025)),
    'D': dict(seed=404, labels_shape=(24, 28, 24), batch=1, forced=[.4], cfg=dict(
        input_channels=[True, False], output_channel=[1], atlas_res=1., target_res=2., output_shape=None, padding_margin=2,
        flipping=True, aff=np.eye(4), scaling_bounds=.15, rotation_bounds=15, shearing_bounds=.012, translation_bounds=False,
        nonlin_std=2., nonlin_shape_factor=.0625, simulate_registration_error=True, randomise_res=False,
        data_res=np.array([[1., 1., 3.]]), thickness=np.array([[1., 1., 3.]]), downsample=False, build_reliability_maps=True,
        blur_range=1.15, bias_field_std=.3, bias_shape_factor=.025)),
    'E': dict(seed=505, labels_shape=(24, 24, 36), batch=1, forced=[.1, .2], cfg=dict(
        input_channels=[True, True], output_channel=[1], atlas_res=1., target_res=1.5, output_shape=None, flipping=True,
        aff=np.eye(4), scaling_bounds=.15, rotation_bounds=15, shearing_bounds=.012, translation_bounds=False, nonlin_std=3.,
        nonlin_shape_factor=.0625, simulate_registration_error=True, randomise_res=False,
        data_res=np.array([[1., 1., 3.], [1., 2.5, 1.]]), thickness=np.array([[1., 1., 2.], [1., 2.5, 1.]]), downsample=True,
        build_reliability_maps=True, blur_range=1.15, bias_field_std=.3, bias_shape_factor=.025)),
    'F': dict(seed=606, labels_shape=(20, 22, 18), batch=2, forced=[.96], cfg=dict(
        input_channels=[True], output_channel=[0], atlas_res=1., target_res=None, output_shape=[16, 16, 16], flipping=False,
        aff=None, scaling_bounds=False, rotation_bounds=20, shearing_bounds=False, translation_bounds=4, nonlin_std=3.,
        nonlin_shape_factor=.125, simulate_registration_error=True, randomise_res=False, data_res=np.array([[2., 1., 1.]]),
        thickness=None, downsample=False, build_reliability_maps=True, blur_range=None, bias_field_std=.2,
        bias_shape_factor=.025)),
    'G': dict(seed=707, labels_shape=(17, 21, 19), batch=1, forced=[.5, .5], cfg=dict(
        input_channels=[False, True, True], output_channel=[0], atlas_res=1., target_res=None, output_shape=None,
        output_div_by_n=8, padding_margin=[2, 0, 3], flipping=True, aff=np.eye(4), scaling_bounds=False, rotation_bounds=False,
        shearing_bounds=False, translation_bounds=False, nonlin_std=0., simulate_registration_error=[True, True, True],
        randomise_res=False, data_res=np.array([[1., 1., 2.], [1., 3., 1.]]), thickness=np.array([[1., 1., 4.], [1., 2., 1.]]),
        downsample=False, build_reliability_maps=False, blur_range=1., bias_field_std=.5, bias_shape_factor=.1)),
    'H': dict(seed=808, labels_shape=(16, 16, 20), batch=1, forced=[.3, .01], cfg=dict(
        input_channels=[True], output_channel=[0], atlas_res=1., target_res=None, output_shape=None, flipping=True,
        aff=np.eye(4), scaling_bounds=.15, rotation_bounds=15, shearing_bounds=.012, translation_bounds=False, nonlin_std=3.,
        nonlin_shape_factor=.0625, simulate_registration_error=True, randomise_res=True, data_res=None, thickness=None,
        downsample=False, build_reliability_maps=True, blur_range=1.15, bias_field_std=.3, bias_shape_factor=.025)),
}

if __name__ == '__main__':
    out, meta = {}, {}
    for tag, c in CASES.items():
        inputs, image, target, log = run(c['seed'], c['labels_shape'], c['batch'], c['forced'], real=c.get('real', False),
                                         **c['cfg'])
        pm = c['cfg'].get('padding_margin') or 0
        pm = (list(np.ravel(pm)) * 3)[:3]
        grid = [s + 2 * int(m) for s, m in zip(c['labels_shape'], pm)]
        # the crop shape the reference derived = shape of the GMM noise it asked for
        crop = [v for k, v, _, _ in log if k == 'normal' and v.ndim == 5][1 if c['cfg'].get('nonlin_std', 3.) > 0 else 0].shape[1:4]
        d = to_draws(c['cfg'], log, c['batch'], list(crop) != grid)
        for i, a in enumerate(inputs):
            out['%s_in%d' % (tag, i)] = a
        out['%s_image' % tag], out['%s_target' % tag] = image, target
        for k, v in d.items():
            if v is not None:
                out['%s_draw_%s' % (tag, k)] = np.asarray(v)
        meta[tag] = dict(labels_shape=list(c['labels_shape']), batch=c['batch'],
                         cfg={k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in c['cfg'].items()},
                         image_shape=list(image.shape), target_shape=list(target.shape),
                         n_draw_calls=len(log), draw_calls=[[k, list(v.shape)] for k, v, _, _ in log],
                         bounds={k: v for k, v in BOUNDS.items() if v is not None})
        print(tag, 'image', image.shape, 'target', target.shape, '%d random ops' % len(log))
    out['generation_labels'] = GEN
    out['n_neutral_labels'] = np.array(N_NEUTRAL)
    out['meta_json'] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, 'reference_model.npz'), **out)
