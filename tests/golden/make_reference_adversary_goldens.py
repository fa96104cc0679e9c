"""Golden vectors for the adversarial fine-tuner (second half): the reference's OWN
SynthSR/fine_tuning_with_adversary.py functions executed on the tf shim --

  build_generator_loss       (:511-577)  with and without loss cropping, with and without the Dice term (layers.DiceLoss executed
                                         from the reference as well)
  build_discriminator_loss   (:580-596)  with the `Gradients` layer (K.gradients, :640) fed a GIVEN gradient array: what is pinned
                                         is the reference's norm (which axes), penalty and sum -- automatic differentiation itself
                                         cannot run on a NumPy shim
  RandomWeightedAverage.call (:619-624)  with the tf.random.uniform draw logged: the shape of the weights and the blend
  make_discriminator         (:482-508)  with recording layer stubs: the layer sequence, filters, strides, units, alphas
                                         (the arithmetic of Conv3D / Dense is restated in oracle/adversary.py, not executed)

Writes tests/golden/reference_adversary.npz (+ reference_adversary_wiring.json).   (build container only: needs /root/reference)"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_model_goldens as MG  # noqa: E402  (tf shim)

shim, T, f32 = MG.shim, MG.T, np.float32
tf = sys.modules['tensorflow']
K = sys.modules['keras.backend']
KL = sys.modules['keras.layers']
f64 = lambda x: np.asarray(x, dtype=np.float64)  # noqa: E731

tf.keras = types.SimpleNamespace(backend=types.SimpleNamespace(epsilon=lambda: 1e-7))
tf.math.reduce_mean = lambda x, axis=None: T(np.mean(f64(x), axis=axis))
tf.math.reduce_sum = lambda x, axis=None, keepdims=False: T(np.sum(f64(x), axis=tuple(axis) if isinstance(axis, list) else axis,
                                                                    keepdims=keepdims))
tf.math.square = lambda x: T(np.square(f64(x)))
tf.stack = lambda xs, axis=0: T(np.stack([np.asarray(x) for x in xs], axis=axis))


def tf_slice(x, begin, size):
    x, begin, size = np.asarray(x), [int(v) for v in np.asarray(begin)], [int(v) for v in np.asarray(size)]
    assert len(begin) == len(size) == x.ndim, 'tf.slice needs one begin / size per axis'      # what TensorFlow enforces
    idx = tuple(slice(b, None if s == -1 else b + s) for b, s in zip(begin, size))
    return T(x[idx])


tf.slice = tf_slice
K.mean = lambda x, axis=None: T(np.mean(f64(x), axis=axis))
K.abs = lambda x: T(np.abs(f64(x)))
K.sqrt = lambda x: T(np.sqrt(f64(x)))
K.sum = lambda x, axis=None: T(np.sum(f64(x), axis=tuple(int(a) for a in np.asarray(axis).reshape(-1)) if axis is not None else None))
K.square = lambda x: T(np.square(f64(x)))
GIVEN_GRADIENTS = []
K.gradients = lambda y, x: [T(GIVEN_GRADIENTS.pop(0))]


class Lambda:
    def __init__(self, fn, name=None, **kw):
        self.fn, self.name = fn, name

    def __call__(self, x):
        return self.fn(x)


KL.Lambda = Lambda
sys.modules['keras.optimizers'].Adam = object
sys.modules['keras'].models = sys.modules['keras.models']
sys.path.insert(0, '/root/reference')
import SynthSR.fine_tuning_with_adversary as RF  # noqa: E402
import ext.lab2im.layers as RL  # noqa: E402

RF.KL = KL
rng = np.random.default_rng(404)
out = {}

# ---------------------------------------------------------------- build_generator_loss
GEN = np.array([0, 1, 2, 3, 4, 5, 14, 15, 41, 42])
# every spatial size > 10: utils.get_dims (ext/lab2im/utils.py:568) takes a last axis <= 10 for a channel axis, and the reference's
# cropping then builds a begin / size of the wrong rank (TensorFlow would refuse it) -- not a case real volumes reach
CASES = {'plain': dict(shape=(12, 11, 13), crop=None, seg=False, dice_w=.25, discr_w=.01),
         'crop': dict(shape=(12, 14, 16), crop=6, seg=False, dice_w=.25, discr_w=.05),
         'seg': dict(shape=(12, 11, 12), crop=None, seg=True, dice_w=.25, discr_w=.01, equiv=np.array([0, 2, 2, 3, -1])),
         'seg_crop': dict(shape=(12, 14, 16), crop=[6, 4, 8], seg=True, dice_w=.3, discr_w=.02, equiv=np.array([0, 1, 1, 1, 4, 5]))}
for name, c in CASES.items():
    shp = c['shape']
    target = rng.uniform(0, 1, size=(2, *shp, 1)).astype(f32)
    pred = rng.uniform(-.1, 1.1, size=(2, *shp, 1)).astype(f32)
    d_out = rng.normal(size=(2, 1)).astype(f32)
    seg_t = seg_o = None
    if c['seg']:
        seg_t = GEN[rng.integers(0, len(GEN), size=(2, *shp, 1))].astype(np.int32)
        z = rng.normal(size=(2, *shp, len(c['equiv'])))
        seg_o = (np.exp(z) / np.exp(z).sum(-1, keepdims=True)).astype(f32)
    shim.base.GRAPH_BATCH[0] = 2
    loss = RF.build_generator_loss(T(target), None if seg_t is None else T(seg_t), T(pred), T(d_out),
                                   None if seg_o is None else T(seg_o), GEN, c.get('equiv'), c['crop'], c['seg'], c['dice_w'],
                                   c['discr_w'])
    shim.base.GRAPH_BATCH[0] = None
    out['gen_%s_target' % name], out['gen_%s_pred' % name], out['gen_%s_dout' % name] = target, pred, d_out
    if c['seg']:
        out['gen_%s_segt' % name], out['gen_%s_sego' % name], out['gen_%s_equiv' % name] = seg_t, seg_o, c['equiv']
    out['gen_%s_loss' % name] = np.array(float(np.asarray(loss, dtype=np.float64).reshape(())))
    print('generator loss', name, out['gen_%s_loss' % name])
out['generation_labels'] = GEN
out['gen_cases'] = np.array(json.dumps({k: {kk: (vv if not isinstance(vv, np.ndarray) else vv.tolist()) for kk, vv in v.items()}
                                        for k, v in CASES.items()}))

# ---------------------------------------------------------------- build_discriminator_loss (gradients given)
for name, (shp, nch, gp_w) in {'a': ((6, 5, 4), 1, 10), 'b': ((4, 4, 6), 2, 3.5)}.items():
    d_real, d_fake, d_av = (rng.normal(size=(3, 1)).astype(f32) for _ in range(3))
    avg = rng.uniform(0, 1, size=(3, *shp, nch)).astype(f32)
    grads = (rng.normal(size=avg.shape) * .2).astype(f32)
    GIVEN_GRADIENTS.append(grads)
    loss = RF.build_discriminator_loss(T(d_real), T(d_fake), T(d_av), T(avg), gp_w, 3)
    for k, v in (('real', d_real), ('fake', d_fake), ('av', d_av), ('samples', avg), ('grads', grads), ('gpw', np.array(gp_w))):
        out['disc_%s_%s' % (name, k)] = v
    out['disc_%s_loss' % name] = np.array(float(np.asarray(loss, dtype=np.float64).reshape(())))
    print('discriminator loss', name, out['disc_%s_loss' % name])

# ---------------------------------------------------------------- RandomWeightedAverage
a, b = rng.uniform(size=(3, 4, 5, 6, 1)).astype(f32), rng.uniform(size=(3, 4, 5, 6, 1)).astype(f32)
draws = []


def logged_uniform(shape, minval=0, maxval=1., dtype='float32'):
    w = rng.uniform(minval, maxval, size=[int(s) for s in np.asarray(shape).reshape(-1)]).astype(f32)
    draws.append(w)
    return T(w)


tf.random.uniform = logged_uniform
layer = RF.RandomWeightedAverage()
layer.build([a.shape, b.shape])
res = layer.call([T(a), T(b)])
out['rwa_a'], out['rwa_b'], out['rwa_w'], out['rwa_out'] = a, b, draws[0], np.asarray(res, dtype=f32)
print('RandomWeightedAverage weights shape', draws[0].shape)

# ---------------------------------------------------------------- make_discriminator wiring
WIRING = []


def rec(kind):
    class L:
        def __init__(self, *args, **kw):
            self.args, self.kw = args, kw

        def __call__(self, x):
            WIRING.append([kind, [a if not isinstance(a, (np.integer, np.floating)) else a.item() for a in self.args],
                           {k: (v if isinstance(v, (int, float, str, type(None), bool)) else str(v)) for k, v in self.kw.items()}])
            return x
    return L


KL.Input = lambda shape=None, name=None, **kw: ('input', name, list(shape))
for kind in ('Conv3D', 'LeakyReLU', 'Flatten', 'Dense'):
    setattr(KL, kind, rec(kind))
RF.models = types.SimpleNamespace(Model=lambda i, o, name=None: types.SimpleNamespace(inputs=i, outputs=o, name=name))
m = RF.make_discriminator([16, 16, 16, 1])
json.dump({'default': WIRING, 'name': m.name}, open(os.path.join(HERE, 'reference_adversary_wiring.json'), 'w'), indent=1)
print(len(WIRING), 'layers in make_discriminator')
np.savez_compressed(os.path.join(HERE, 'reference_adversary.npz'), **out)
