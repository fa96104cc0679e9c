"""Golden vectors for the host-side input sampler: the reference's own SynthSR/model_inputs.build_model_inputs (NumPy only)
run on .npz label maps with the global NumPy generator seeded; the test reseeds and runs the product's sampler.
Writes tests/golden/reference_model_inputs.npz (+ the tiny label maps it used).   (build container only)"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_numpy_shim  # noqa: E402

tf_numpy_shim.install([])
sys.path.insert(0, '/root/reference')
from SynthSR.model_inputs import build_model_inputs  # noqa: E402

out = {}
rng = np.random.default_rng(31)
K = 7
maps = [rng.integers(0, K, size=(6, 5, 4)).astype(np.int32) for _ in range(3)]
imgs = [rng.uniform(0, 255, size=(6, 5, 4)).astype(np.float32) for _ in range(3)]
tmp = tempfile.mkdtemp()
lp, ip = [], []
for i, (m, im) in enumerate(zip(maps, imgs)):
    lp.append(os.path.join(tmp, 'lab%d.npz' % i))
    ip.append(os.path.join(tmp, 'img%d.npz' % i))
    np.savez(lp[-1], vol_data=m)
    np.savez(ip[-1], vol_data=im)
    out['map_%d' % i], out['img_%d' % i] = m, im

classes = np.array([0, 1, 2, 2, 3, 3, 4])
pm2 = np.stack([rng.uniform(20, 200, size=5), rng.uniform(1, 30, size=5)])          # (2, n_classes)
ps2 = np.stack([rng.uniform(5, 20, size=5), rng.uniform(1, 5, size=5)])
pm4 = np.concatenate([np.sort(rng.uniform(20, 220, size=(2, K)), 0), np.sort(rng.uniform(20, 220, size=(2, K)), 0)])
ps4 = np.concatenate([np.sort(rng.uniform(2, 30, size=(2, K)), 0), np.sort(rng.uniform(2, 30, size=(2, K)), 0)])
out.update(classes=classes, pm2=pm2, ps2=ps2, pm4=pm4, ps4=ps4)
CASES = {
    'default': dict(n_labels=K, prior_means=None, prior_stds=None, prior_distributions='uniform'),
    'normal_classes': dict(n_labels=K, prior_means=pm2, prior_stds=ps2, prior_distributions='normal', generation_classes=classes),
    'two_channels_images': dict(n_labels=K, prior_means=pm4, prior_stds=ps4, prior_distributions='uniform', n_channels=2,
                                batchsize=2, path_images=ip),
    'range_pair': dict(n_labels=K, prior_means=[40, 180], prior_stds=[3, 12], prior_distributions='uniform'),
}
for name, kw in CASES.items():
    np.random.seed(1234)
    g = build_model_inputs(lp, **kw)
    for it in range(3):
        res = next(g)
        for j, a in enumerate(res):
            out['%s_it%d_%d' % (name, it, j)] = np.asarray(a)
    print(name, [np.asarray(a).shape for a in res])
np.savez_compressed(os.path.join(HERE, 'reference_model_inputs.npz'), **out)
