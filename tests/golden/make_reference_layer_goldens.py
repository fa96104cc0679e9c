"""Golden vectors for the reference's intensity / label LAYERS from the reference's OWN code executed on the NumPy `tf`
shim (tf_numpy_shim_layers.py): ext/lab2im/layers.py SampleConditionalGMM, BiasFieldCorruption, IntensityAugmentation,
RandomFlip, RandomCrop, GaussianBlur -- every tf.random draw injected from a queue and saved with the outputs, so the
oracle (oracle/generator.py) can be run on the same draws.  Writes tests/golden/reference_layers.npz.
(build container only: needs /root/reference)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_numpy_shim_layers  # noqa: E402

queue = []
tf, K, T = tf_numpy_shim_layers.install(queue)
sys.path.insert(0, '/root/reference')
from ext.lab2im import layers  # noqa: E402

f32 = np.float32
rng = np.random.default_rng(11)
out = {}
GEN = np.array([0, 14, 15, 16, 2, 3, 4, 5, 7, 8, 10, 11, 12, 13, 17, 18, 26, 28, 31])      # data/labels_classes_priors


def run(layer, inputs, in_shapes):
    layer.build(in_shapes)
    assert not queue or True
    return layer.call(inputs)


# ---- SampleConditionalGMM (layers.py:430-501), batch 1 and batch 2, two channels -------------------------------------
for B in (1, 2):
    shape = (B, 6, 5, 7)
    labels = GEN[rng.integers(0, len(GEN), size=shape)].astype(np.int32)[..., None]
    means = rng.uniform(0, 250, size=(B, len(GEN), 2)).astype(f32)
    stds = rng.uniform(0, 30, size=(B, len(GEN), 2)).astype(f32)
    noise = rng.normal(size=shape + (2,)).astype(f32)
    queue.append(noise)
    y = run(layers.SampleConditionalGMM(GEN), [T(labels), T(means), T(stds)],
            [(None,) + shape[1:] + (1,), (None, len(GEN), 2), (None, len(GEN), 2)])
    assert not queue
    out.update({'gmm%d_labels' % B: labels, 'gmm%d_means' % B: means, 'gmm%d_stds' % B: stds, 'gmm%d_noise' % B: noise,
                'gmm%d_out' % B: np.asarray(y)})

# ---- BiasFieldCorruption (layers.py:1002-1097): applied (u < .95) and skipped (u >= .95) ------------------------------
for tag, u in (('on', .3), ('off', .97)):
    shape = (1, 20, 24, 16, 1)
    x = rng.uniform(0, 300, size=shape).astype(f32)
    layer = layers.BiasFieldCorruption(.3, .2, False)
    layer.build(shape and (None,) + shape[1:])
    small = [int(v) for v in layer.small_bias_shape]
    std = rng.uniform(0, .3, size=(1, 1, 1, 1, 1)).astype(f32)
    nrm = rng.normal(size=(1, *small)).astype(f32)
    queue.extend([std, nrm, np.array([u], dtype=f32)])                  # evaluation order: uniform(std), normal, uniform(prob)
    y = layer.call(T(x))
    assert not queue
    out.update({'bias_%s_x' % tag: x, 'bias_%s_std' % tag: std, 'bias_%s_normal' % tag: nrm, 'bias_%s_out' % tag: np.asarray(y),
                'bias_small_shape': np.array(small)})

# ---- IntensityAugmentation(clip=300, normalise=True, gamma_std=.5, separate_channels=True) (labels_to_image_model.py:184)
shape = (2, 9, 8, 10, 1)
x = rng.uniform(-40, 420, size=shape).astype(f32)
gamma = rng.normal(size=(2, 1, 1, 1, 1)).astype(f32)
queue.append(gamma)
layer = layers.IntensityAugmentation(clip=300, normalise=True, gamma_std=.5, separate_channels=True)
layer.build((None,) + shape[1:])
y = layer.call(T(x))
assert not queue
out.update({'int_x': x, 'int_gamma': gamma, 'int_out': np.asarray(y)})

# ---- RandomFlip(flip_axis 0, swap_labels, label list with 3 neutral + 2x2 sided labels) (labels_to_image_model.py:159-162)
label_list = np.array([0, 14, 15, 2, 3, 41, 42])
n_neutral = 3
shape = (3, 6, 4, 5, 1)
lab = label_list[rng.integers(0, len(label_list), size=shape)].astype(np.int32)
img = rng.uniform(size=shape).astype(f32)
u = np.array([[.2], [.7], [.49]], dtype=f32)                           # flip, no flip, flip   (prob .5)
queue.append(u)
layer = layers.RandomFlip(0, [True, False], label_list, n_neutral)
layer.build([(None,) + shape[1:], (None,) + shape[1:]])
yl, yi = layer.call([T(lab), T(img)])
assert not queue
out.update({'flip_labels': lab, 'flip_image': img, 'flip_u': u, 'flip_label_list': label_list, 'flip_out_labels': np.asarray(yl),
            'flip_out_image': np.asarray(yi)})

# ---- RandomCrop (layers.py:214-274): same offsets for both inputs ----------------------------------------------------
shape = (2, 11, 9, 12, 1)
crop = [8, 9, 7]
a = rng.integers(0, 50, size=shape).astype(np.int32)
b = rng.uniform(size=shape).astype(f32)
maxv = np.array(shape[1:4]) - np.array(crop)
u = (rng.uniform(size=(2, 3)) * maxv).astype(f32)
queue.extend([u[0], u[1]])
layer = layers.RandomCrop(crop)
layer.build([(None,) + shape[1:], (None,) + shape[1:]])
ya, yb = layer.call([T(a), T(b)])
assert not queue
out.update({'crop_a': a, 'crop_b': b, 'crop_u': u, 'crop_shape': np.array(crop), 'crop_out_a': np.asarray(ya),
            'crop_out_b': np.asarray(yb)})

# ---- GaussianBlur(sigma=.5) and GaussianBlur(sigma=.42*[1,1,3], random_blur_range=1.15) (labels_to_image_model.py:186, 224)
shape = (1, 10, 9, 12, 1)
x = rng.uniform(size=shape).astype(f32)
layer = layers.GaussianBlur(.5)
layer.build((None,) + shape[1:])
y = layer.call(T(x))
out.update({'blur_x': x, 'blur_out_05': np.asarray(y)})
mult = rng.uniform(1 / 1.15, 1.15, size=(3,)).astype(f32)
queue.append(mult)
sig = [.42, .42, 1.26]
layer = layers.GaussianBlur(sig, 1.15)
layer.build((None,) + shape[1:])
y = layer.call(T(x))
assert not queue
out.update({'blur_mult': mult, 'blur_sigma': np.array(sig), 'blur_out_acq': np.asarray(y)})

np.savez_compressed(os.path.join(HERE, 'reference_layers.npz'), **out)
print({k: v.shape for k, v in out.items() if 'out' in k})
