"""Golden vectors for the training loss: the reference's own SynthSR/metrics_model.metrics_model() executed on the NumPy
`tf` shim on top of a stand-in input model (prediction, 'image_out' and 'regression_target' layer outputs fed as arrays):
residual addition (work_with_residual_channel), centre cropping (loss_cropping), L1 / L2 reduction.
Writes tests/golden/reference_loss.npz.   (build container only: needs /root/reference)"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_numpy_shim_layers as shim  # noqa: E402

tf, K, T = shim.install([])
f32 = np.float32
KL = sys.modules['keras.layers']


class _Merge:
    def __init__(self, name=None, **kw):
        pass


class Add(_Merge):
    def __call__(self, xs):
        return T((np.asarray(xs[0]) + np.asarray(xs[1])).astype(f32))


class Subtract(_Merge):
    def __call__(self, xs):
        return T((np.asarray(xs[0]) - np.asarray(xs[1])).astype(f32))


KL.Add, KL.Subtract = Add, Subtract
K.mean = lambda x, axis=None: T(np.mean(np.asarray(x), axis=axis, dtype=np.float64).astype(f32))   # order-free reference value
K.abs = lambda x: T(np.abs(np.asarray(x)))


class Model:
    def __init__(self, inputs=None, outputs=None):
        self.inputs, self.outputs = inputs, outputs


sys.modules['keras.models'].Model = Model
sys.path.insert(0, '/root/reference')
from SynthSR.metrics_model import metrics_model  # noqa: E402


class FakeInputModel:
    def __init__(self, pred, image, target):
        self.inputs = []
        self.outputs = [T(pred)]
        self._layers = {'image_out': types.SimpleNamespace(output=T(image)),
                        'regression_target': types.SimpleNamespace(output=T(target))}

    def get_layer(self, name):
        return self._layers[name]


rng = np.random.default_rng(23)
out = {}
CASES = [
    ('l1_plain', (1, 12, 10, 14), 2, 1, dict(metrics='l1', loss_cropping=None, work_with_residual_channel=None)),
    ('l2_crop_res', (2, 16, 16, 12), 2, 1, dict(metrics='l2', loss_cropping=8, work_with_residual_channel=[0])),
    ('l1_crop3_res2', (1, 14, 18, 16), 4, 2, dict(metrics='l1', loss_cropping=[8, 10, 6], work_with_residual_channel=[0, 2])),
    ('l1_crop_odd', (1, 15, 13, 17), 1, 1, dict(metrics='l1', loss_cropping=6, work_with_residual_channel=None)),
]
for name, shp, cimg, cout, kw in CASES:
    pred = rng.normal(size=shp + (cout,)).astype(f32)
    image = rng.uniform(size=shp + (cimg,)).astype(f32)
    target = rng.uniform(size=shp + (cout,)).astype(f32)
    shim.base.GRAPH_BATCH[0] = shp[0]
    m = metrics_model(FakeInputModel(pred, image, target), **kw)
    shim.base.GRAPH_BATCH[0] = None
    out[name + '_pred'], out[name + '_image'], out[name + '_target'] = pred, image, target
    out[name + '_loss'] = np.asarray(m.outputs, dtype=f32).reshape(())
    print(name, float(out[name + '_loss']))
np.savez_compressed(os.path.join(HERE, 'reference_loss.npz'), **out)
