"""Golden vectors for the randomise_res branch from the reference's OWN code executed on the NumPy `tf` shim:
ext/lab2im/edit_tensors.py blurring_sigma_for_downsampling (tensor branch) and gaussian_kernel (sigma given as a
[B,3] tensor, separable), ext/lab2im/layers.py MimicAcquisition.call (nearest down, linear up, distance map).
Writes tests/golden/reference_randomise_res.npz.   (build container only: needs /root/reference)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_numpy_shim  # noqa: E402

queue = []
tf, K, T = tf_numpy_shim.install(queue)
sys.path.insert(0, '/root/reference')
from ext.lab2im import edit_tensors as et  # noqa: E402
from ext.lab2im import layers  # noqa: E402

f32 = np.float32
rng = np.random.default_rng(7)
out = {}


def batched(a):
    t = T(np.asarray(a, dtype=f32))
    t.static_batch_unknown = True          # Keras tensors have batch dimension None when the graph is built
    return t


B = 3
atlas_res = np.array([1., 1., 1.])
res = rng.uniform(1., 9., size=(B, 3)).astype(f32)
res[1] = [1., 1., 1.]                                              # one example at the atlas resolution
thick = (1. + rng.uniform(size=(B, 3)) * (res - 1.)).astype(f32)
out['res'], out['thick'] = res, thick

# sigma = blurring_sigma_for_downsampling(atlas_res, resolution, mult_coef=.42, thickness=blur_res)   (labels_to_image_model.py:218)
sigma = et.blurring_sigma_for_downsampling(atlas_res, batched(res), mult_coef=.42, thickness=batched(thick))
out['sigma'] = np.asarray(sigma)

# kernels of DynamicGaussianBlur(0.75 * max_res / atlas_res, blur_range=1.15)   (layers.py:813)
max_sigma = 0.75 * np.array([9.] * 3) / atlas_res
mult = rng.uniform(1 / 1.15, 1.15, size=(B, 3)).astype(f32)
out['mult'] = mult
queue.append(mult)
ks = et.gaussian_kernel(batched(np.asarray(sigma)), max_sigma, 1.15, True)
for i, k in enumerate(ks):
    out['kernel_%d' % i] = np.asarray(k).reshape(B, -1)
ks0 = et.gaussian_kernel(batched(np.asarray(sigma)), max_sigma, None, True)
out['kernel_nojitter_2'] = np.asarray(ks0[2]).reshape(B, -1)

# MimicAcquisition(atlas_res, atlas_res, output_shape, build_dist_map=True)([channel, resolution])   (:220)
inshape = [12, 10, 14]
resample = [8, 10, 16]
vol = rng.uniform(size=(B, *inshape, 1)).astype(f32)
out['vol'] = vol
layer = layers.MimicAcquisition(atlas_res, atlas_res, resample, True)
layer.build([(None, *inshape, 1), (None, 3)])
res_m = np.array([[1., 1., 1.], [1.5, 2.2, 3.], [4.1, 1., 8.9]], dtype=f32)
out['res_mimic'] = res_m
o, d = layer.call([T(vol), T(res_m)])
out['mimic_vol'], out['mimic_dist'] = np.asarray(o), np.asarray(d)

np.savez_compressed(os.path.join(HERE, 'reference_randomise_res.npz'), **out)
print({k: (v.shape, v.dtype) for k, v in out.items()})
