"""Golden vectors for the segmentation-regularised loss (NOT built on the GPU yet -- this pins the
oracle ahead of the kernels): the reference's own SynthSR/metrics_model.add_seg_loss_to_model() executed on the tf shim, the
frozen segmentation network being the reference's own ext.neuron.models.unet(final_pred_activation='softmax') built by its
builder on the functional Keras stand-in, and layers.DiceLoss(enable_checks=False) executed from the reference.

Keras semantics that are restated, not executed (TF / Keras are not installable): the softmax, and BatchNormalization of
the frozen (trainable=False) network normalising with BATCH statistics while fitting (Keras 2.3.1's BatchNormalization.call
does not look at `trainable`; only the moving-average updates are dropped).

Writes tests/golden/reference_segloss.npz.   (build container only: needs /root/reference)"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_model_goldens as MG  # noqa: E402  (tf shim)
import make_reference_unet_goldens as UG  # noqa: E402   (functional Keras stand-in)

shim, T, f32 = MG.shim, MG.T, np.float32
tf = sys.modules['tensorflow']
K = sys.modules['keras.backend']
KL = sys.modules['keras.layers']
UG.install()


def softmax(x, axis=-1):
    x = np.asarray(x, dtype=np.float64)
    e = np.exp(x - x.max(axis=axis, keepdims=True))
    return e / e.sum(axis=axis, keepdims=True)


sys.modules['keras'].activations = types.SimpleNamespace(softmax=softmax)
sys.modules['keras'].layers = KL
tf.keras = types.SimpleNamespace(backend=types.SimpleNamespace(epsilon=lambda: 1e-7))
tf.math.reduce_mean = lambda x, axis=None: T(np.mean(np.asarray(x, dtype=np.float64), axis=axis))
tf.math.reduce_sum = lambda x, axis=None, keepdims=False: T(np.sum(np.asarray(x, dtype=np.float64),
                                                                    axis=tuple(axis) if isinstance(axis, list) else axis,
                                                                    keepdims=keepdims))
tf.math.square = lambda x: T(np.square(np.asarray(x, dtype=np.float64)))
tf.stack = lambda xs, axis=0: T(np.stack([np.asarray(x) for x in xs], axis=axis))
K.reverse = lambda x, axes: T(np.flip(np.asarray(x), axis=axes))
K.clip = lambda x, lo, hi: T(np.clip(np.asarray(x), lo, hi))


class Lambda:
    def __init__(self, fn, name=None, **kw):
        self.fn, self.name = fn, name

    def __call__(self, x):
        self.output = self.fn(x)
        return self.output


KL.Lambda = Lambda
sys.path.insert(0, '/root/reference')
import SynthSR.metrics_model as RM  # noqa: E402
import ext.neuron.models as nrn_models  # noqa: E402

RM.KL = KL
nrn_models.keras = sys.modules['keras']


class Model:
    def __init__(self, inputs=None, outputs=None, name=None):
        self.inputs, self.outputs = inputs, outputs


RM.Model = Model
SEG = dict(nb_features=4, nb_levels=3, conv_size=3, feat_mult=2, nb_conv_per_level=2)


def make_seg_model(n_seg_labels, weights):
    """`seg_model(tensor)`: the reference's unet() builder run on the tensor it is called with (weights by Keras name)."""
    def call(x):
        saved = (dict(UG.LAYERS), dict(UG.WEIGHTS))
        UG.LAYERS.clear()
        UG.WEIGHTS.clear()
        UG.WEIGHTS.update(weights)
        UG.FEED.append(UG.KTensor(np.asarray(x, dtype=np.float64)))
        m = nrn_models.unet(input_shape=list(np.asarray(x).shape[1:]), nb_labels=n_seg_labels, conv_dropout=0,
                            final_pred_activation='softmax', batch_norm=-1, activation='elu', input_model=None, **SEG)
        weights.update(UG.WEIGHTS)                         # created on first use with the shapes the builder asked for
        out = np.asarray(m.output, dtype=np.float64)
        UG.LAYERS.clear(); UG.LAYERS.update(saved[0])
        UG.WEIGHTS.clear(); UG.WEIGHTS.update(saved[1])
        return T(out)
    return call


class FakeInputModel:
    def __init__(self, image_loss, predicted_image, segm_target):
        self.inputs = []
        self.outputs = [T(np.asarray(image_loss, dtype=np.float64))]
        self._layers = {'predicted_image': types.SimpleNamespace(output=T(predicted_image)),
                        'segmentation_target': types.SimpleNamespace(output=T(segm_target))}

    def get_layer(self, name):
        return self._layers[name]


rng = np.random.default_rng(91)
GEN = np.array([0, 1, 2, 3, 4, 5, 14, 15, 41, 42])          # label VALUES; note the reference compares them to loop INDICES
out = {'generation_labels': GEN}
CASES = {
    # segmentation labels (5 outputs) -> generation labels; -1 = ignored; 2 and 41 merged onto... see equivalency
    'plain': dict(shape=(16, 16, 16), equiv=np.array([0, 2, 2, 3, -1]), rel_weight=.25, loss_cropping=None, m=None, M=None,
                  fs_header=False),
    'crop_clip_fs': dict(shape=(16, 12, 16), equiv=np.array([0, 1, 1, 1, 4, 5]), rel_weight=.5, loss_cropping=8, m=.1, M=.8,
                         fs_header=True),
}
for name, c in CASES.items():
    shp = c['shape']
    pred_img = rng.uniform(-.1, 1.1, size=(1, *shp, 1)).astype(f32)
    seg_t = GEN[rng.integers(0, len(GEN), size=(1, *shp, 1))].astype(np.int32)
    image_loss = f32(rng.uniform(.05, .2))
    weights = {}
    shim.base.GRAPH_BATCH[0] = 1
    model = RM.add_seg_loss_to_model(FakeInputModel(image_loss, pred_img, seg_t), make_seg_model(len(c['equiv']), weights),
                                     GEN, c['equiv'], c['rel_weight'], c['loss_cropping'], m=c['m'], M=c['M'],
                                     fs_header=c['fs_header'])
    shim.base.GRAPH_BATCH[0] = None
    total = float(np.asarray(model.outputs, dtype=np.float64).reshape(()))
    out[name + '_pred_image'], out[name + '_seg_target'], out[name + '_image_loss'] = pred_img, seg_t, np.array(image_loss)
    out[name + '_equiv'] = c['equiv']
    out[name + '_total'] = np.array(total)
    for k, v in weights.items():
        out['%s_w/%s' % (name, k)] = v
    print(name, 'total', total, 'dice', (total - float(image_loss)) / c['rel_weight'], len(weights), 'weights')
np.savez_compressed(os.path.join(HERE, 'reference_segloss.npz'), **out)
