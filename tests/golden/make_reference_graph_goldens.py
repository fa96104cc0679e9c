"""Golden vectors from the reference's OWN graph functions, executed here on the NumPy `tf` shim
(tests/golden/tf_numpy_shim.py): ext/neuron/utils.py interpn / resize / transform / integrate_vec / affine_to_shift /
combine_non_linear_and_aff_to_shift and ext/lab2im/edit_tensors.py gaussian_kernel / blurring_sigma_for_downsampling.
Writes tests/golden/reference_graph_ops.npz.   (build container only: needs /root/reference)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_numpy_shim  # noqa: E402

queue = []
tf, K, T = tf_numpy_shim.install(queue)
sys.path.insert(0, '/root/reference')
from ext.neuron import utils as nu  # noqa: E402
from ext.lab2im import edit_tensors as et  # noqa: E402

f32 = np.float32
rng = np.random.default_rng(42)
out = {}

vol = rng.normal(size=(6, 7, 5, 3)).astype(f32)
loc = [rng.uniform(-2, s + 1, size=(4, 5, 6)).astype(f32) for s in (6, 7, 5)]
loc[0][0, 0, :4] = [0.5, 1.5, 2.5, 3.5]                       # exact halves: exercises round-half-even
out['vol'], out['loc'] = vol, np.stack(loc, -1)
out['interpn_linear'] = np.asarray(nu.interpn(T(vol), [T(l) for l in loc], 'linear'))
out['interpn_nearest'] = np.asarray(nu.interpn(T(vol), [T(l) for l in loc], 'nearest'))

small = rng.normal(size=(3, 4, 2, 3)).astype(f32)
out['small'] = small
zoom = [8 / 3, 9 / 4, 5 / 2]
out['resize_linear'] = np.asarray(nu.resize(T(small), zoom, [8, 9, 5], 'linear'))
out['resize_nearest'] = np.asarray(nu.resize(T(small), zoom, [8, 9, 5], 'nearest'))
down = [3 / 6, 4 / 7, 2 / 5]
out['resize_down_nearest'] = np.asarray(nu.resize(T(vol), down, [3, 4, 2], 'nearest'))

field = (rng.normal(size=(8, 9, 5, 3)) * 1.5).astype(f32)
out['field'] = field
out['transform_linear'] = np.asarray(nu.transform(T(out['resize_linear']), T(field), 'linear'))
out['integrate_vec'] = np.asarray(nu.integrate_vec(T(field), method='ss', nb_steps=7))

aff = np.eye(4, dtype=f32)
aff[:3, :3] += rng.normal(size=(3, 3)).astype(f32) * f32(0.1)
aff[:3, 3] = [1.5, -2.25, 0.75]
out['aff'] = aff
out['affine_to_shift'] = np.asarray(nu.affine_to_shift(T(aff), [8, 9, 5], shift_center=True))
out['combine_shift'] = np.asarray(nu.combine_non_linear_and_aff_to_shift([T(field), T(aff)], [8, 9, 5], shift_center=True))
labels = rng.integers(0, 30, size=(8, 9, 5, 1)).astype(f32)
out['labels'] = labels
out['labels_warped'] = np.asarray(nu.transform(T(labels), T(out['combine_shift']), 'nearest'))

# gaussian kernels (dense, dense with jitter, separable)
out['gk_05'] = np.asarray(et.gaussian_kernel([.5, .5, .5], separable=False))[..., 0, 0]
mult = rng.uniform(1 / 1.15, 1.15, size=3).astype(f32)
out['gk_mult'] = mult
queue.append(mult)
out['gk_acq_jitter'] = np.asarray(et.gaussian_kernel([.42, .42, 1.26], blur_range=1.15, separable=False))[..., 0, 0]
out['gk_zero_axis'] = np.asarray(et.gaussian_kernel([.5, 0., .75], separable=False))[..., 0, 0]
ks = et.gaussian_kernel([6., .3, 2.1], separable=True)
out['gk_sep_0'] = np.asarray(ks[0]).reshape(-1)
out['gk_sep_2'] = np.asarray(ks[2]).reshape(-1)
assert ks[1] is None

np.savez_compressed(os.path.join(HERE, 'reference_graph_ops.npz'), **out)
print({k: v.shape for k, v in out.items()})
