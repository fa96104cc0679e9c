"""Golden vectors for the spatial augmentation from the reference's OWN code executed on the NumPy `tf` shim:
ext/lab2im/utils.py sample_affine_transform (+ create_rotation_transform / create_shearing_transform /
draw_value_from_distribution) and the complete ext/lab2im/layers.py RandomSpatialDeformation layer (affine + SVF ->
Resize -> VecInt -> Resize -> SpatialTransformer, labels 'nearest' and an image 'linear' warped by the same transforms),
every tf.random draw injected and saved.  Writes tests/golden/reference_spatial.npz.
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
from ext.lab2im import layers, utils  # noqa: E402

f32 = np.float32
rng = np.random.default_rng(17)
out = {}


def affine_draws(B, rot=15., sc=.15, sh=.02, tr=5.):
    return [rng.uniform(-rot, rot, size=(B, 3)).astype(f32), rng.uniform(-sh, sh, size=(B, 6)).astype(f32),
            rng.uniform(1 - sc, 1 + sc, size=(B, 3)).astype(f32), rng.uniform(-tr, tr, size=(B, 3)).astype(f32)]


# ---- sample_affine_transform with the training() defaults (training.py:59-62), batch 3 -----------------------------
d = affine_draws(3)
queue.extend(d)                                             # evaluation order: rotation, shearing, scaling, translation
A = utils.sample_affine_transform(T(np.array([3], np.int32)), 3, 15, .15, .02, 5)
assert not queue
out.update({'aff_rotation': d[0], 'aff_shearing': d[1], 'aff_scaling': d[2], 'aff_translation': d[3], 'aff_out': np.asarray(A)})

# ---- RandomSpatialDeformation(training defaults, nonlin_std 4, factor .0625), labels (nearest) + image (linear) --------
shape = (1, 20, 24, 18, 1)
lab = rng.integers(0, 30, size=shape).astype(np.int32)
img = rng.uniform(0, 1, size=shape).astype(f32)
layer = layers.RandomSpatialDeformation(scaling_bounds=.15, rotation_bounds=15, shearing_bounds=.02, translation_bounds=5,
                                        nonlin_std=4., nonlin_shape_factor=.0625, inter_method=['nearest', 'linear'])
layer.build([(None,) + shape[1:], (None,) + shape[1:]])
small = [int(v) for v in layer.small_shape]
d = affine_draws(1)
std = rng.uniform(0, 4., size=(1, 1)).astype(f32)
svf = rng.normal(size=(1, *small)).astype(f32)
queue.extend(d + [std, svf])
yl, yi = layer.call([T(lab), T(img)])
assert not queue
out.update({'rsd_labels': lab, 'rsd_image': img, 'rsd_rotation': d[0], 'rsd_shearing': d[1], 'rsd_scaling': d[2],
            'rsd_translation': d[3], 'rsd_svf_std': std, 'rsd_svf_normal': svf, 'rsd_small_shape': np.array(small),
            'rsd_out_labels': np.asarray(yl), 'rsd_out_image': np.asarray(yi)})

# ---- same layer, elastic only (no affine) and affine only (no elastic) ---------------------------------------------
layer = layers.RandomSpatialDeformation(scaling_bounds=False, rotation_bounds=False, shearing_bounds=False,
                                        translation_bounds=False, nonlin_std=3., nonlin_shape_factor=.0625,
                                        inter_method='nearest')
layer.build((None,) + shape[1:])
std = rng.uniform(0, 3., size=(1, 1)).astype(f32)
svf = rng.normal(size=(1, *[int(v) for v in layer.small_shape])).astype(f32)
queue.extend([std, svf])
y = layer.call(T(lab))
assert not queue
out.update({'el_svf_std': std, 'el_svf_normal': svf, 'el_out_labels': np.asarray(y)})
layer = layers.RandomSpatialDeformation(scaling_bounds=.15, rotation_bounds=15, shearing_bounds=.02, translation_bounds=False,
                                        nonlin_std=0., inter_method='nearest')
layer.build((None,) + shape[1:])
d = affine_draws(1)[:3]
queue.extend(d)
y = layer.call(T(lab))
assert not queue
out.update({'af_rotation': d[0], 'af_shearing': d[1], 'af_scaling': d[2], 'af_out_labels': np.asarray(y)})

np.savez_compressed(os.path.join(HERE, 'reference_spatial.npz'), **out)
print({k: v.shape for k, v in out.items() if 'out' in k})
