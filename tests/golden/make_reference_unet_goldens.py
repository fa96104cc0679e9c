"""Golden vectors for the U-Net WIRING: the reference's own ext/neuron/models.unet() (-> conv_enc, conv_dec) is executed,
unmodified, on a functional-API stand-in for Keras: KL.Input returns the fed array; Conv3D / BatchNormalization /
MaxPooling3D / UpSampling3D / concatenate / Activation are evaluated eagerly in float64 from their published definitions
(cross-correlation with 'same' zero padding + bias + activation; training-mode BN with the biased batch variance and
epsilon 1e-3; 2x2x2 max-pool with 'same' padding; nearest up-sampling), with the layer weights looked up BY THE NAME the
reference gives each layer.  What this pins is everything the reference's builder decides: layer names and order, feature
counts, kernel sizes, which activation sits where, where BatchNorm is applied, which tensor the skip connection takes
(`get_layer(conv_downarm_l_1).output`), the concatenation order, the final 1x1x1 'likelihood' convolution and the linear
prediction.  The layer arithmetic itself is a restatement (TF is not installable here) -- see DESIGN.md section 2.

Writes tests/golden/reference_unet.npz: weights (by Keras name), input, prediction, and a few intermediate activations.
(build container only: needs /root/reference)"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_numpy_shim  # noqa: E402

if 'tensorflow' not in sys.modules:                   # (imported by make_reference_training_goldens.py after the tf shim is up)
    tf_numpy_shim.install([])
KL = sys.modules['keras.layers']

WEIGHTS = {}          # '<layer>/<weight>' -> array, created on first use with the shapes the reference's layers ask for
LAYERS = {}           # name -> layer object (with .output), in creation order
RNG = np.random.default_rng(77)
FEED = []


class KTensor(np.ndarray):
    def __new__(cls, a):
        return np.asarray(a, dtype=np.float64).view(cls)

    def get_shape(self):
        return tf_numpy_shim.TensorShape((None,) + tuple(np.ndarray.shape.__get__(self)[1:]))

    @property
    def shape(self):
        return tf_numpy_shim.TensorShape((None,) + tuple(np.ndarray.shape.__get__(self)[1:]))


def arr(x):
    return np.asarray(x, dtype=np.float64)


def act(x, name):
    if name in (None, 'linear'):
        return x
    if name == 'elu':
        return np.where(x > 0, x, np.expm1(np.minimum(x, 0)))
    raise NotImplementedError(name)


class _L:
    def __init__(self, name=None):
        assert name is not None and name not in LAYERS, name
        self.name = name
        LAYERS[name] = self

    def __call__(self, x):
        self.input = x
        self.output = KTensor(self.call(x))
        return self.output


class Conv3D(_L):
    def __init__(self, filters, kernel_size, padding='valid', activation=None, data_format='channels_last', dilation_rate=1,
                 name=None):
        super().__init__(name)
        self.filters, self.k, self.activation = int(filters), int(kernel_size), activation
        assert (padding == 'same' or self.k == 1) and data_format == 'channels_last' and int(dilation_rate) == 1   # 'valid' 1x1x1 == 'same'

    def call(self, x):
        x = arr(x)
        cin = x.shape[-1]
        kk, bb = self.name + '/kernel', self.name + '/bias'
        if kk not in WEIGHTS:
            fan = self.k ** 3 * cin
            WEIGHTS[kk] = (RNG.normal(size=(self.k,) * 3 + (cin, self.filters)) * np.sqrt(2. / fan)).astype(np.float32)
            WEIGHTS[bb] = RNG.normal(size=self.filters).astype(np.float32) * np.float32(.1)
        w = torch.from_numpy(WEIGHTS[kk].astype(np.float64)).permute(4, 3, 0, 1, 2)
        y = F.conv3d(torch.from_numpy(x).permute(0, 4, 1, 2, 3), w, torch.from_numpy(WEIGHTS[bb].astype(np.float64)),
                     padding=self.k // 2)
        return act(y.permute(0, 2, 3, 4, 1).numpy(), self.activation)


class BatchNormalization(_L):
    def __init__(self, axis=-1, name=None):
        super().__init__(name)
        assert axis == -1
        self.eps = 1e-3                                                        # Keras default

    def call(self, x):
        x = arr(x)
        c = x.shape[-1]
        for w, init in (('gamma', lambda: RNG.uniform(.5, 1.5, size=c)), ('beta', lambda: RNG.normal(size=c) * .1)):
            if self.name + '/' + w not in WEIGHTS:
                WEIGHTS[self.name + '/' + w] = init().astype(np.float32)
        mean = x.mean(axis=(0, 1, 2, 3))
        var = x.var(axis=(0, 1, 2, 3))                                         # biased, training mode
        g, b = WEIGHTS[self.name + '/gamma'].astype(np.float64), WEIGHTS[self.name + '/beta'].astype(np.float64)
        return (x - mean) / np.sqrt(var + self.eps) * g + b


class MaxPooling3D(_L):
    def __init__(self, pool_size=2, name=None, padding='valid'):
        super().__init__(name)
        assert tuple(pool_size) == (2, 2, 2) and padding == 'same'

    def call(self, x):
        x = arr(x)
        pads = [(0, s % 2) for s in x.shape[1:4]]                              # 'same': pad at the end, -inf
        x = np.pad(x, [(0, 0)] + pads + [(0, 0)], constant_values=-np.inf)
        B, X, Y, Z, C = x.shape
        return x.reshape(B, X // 2, 2, Y // 2, 2, Z // 2, 2, C).max(axis=(2, 4, 6))


class UpSampling3D(_L):
    def __init__(self, size=2, name=None):
        super().__init__(name)
        assert tuple(size) == (2, 2, 2)

    def call(self, x):
        return arr(x).repeat(2, 1).repeat(2, 2).repeat(2, 3)


class Activation(_L):
    def __init__(self, activation, name=None):
        super().__init__(name)
        self.activation = activation

    def call(self, x):
        return act(arr(x), self.activation)


class _Input(_L):
    def call(self, x):
        return x


def Input(shape=None, name=None, dtype=None):
    return _Input(name)(FEED.pop(0))


def concatenate(xs, axis=-1, name=None):
    class _C(_L):
        def call(self, x):
            return np.concatenate([arr(v) for v in x], axis=axis)
    return _C(name)(xs)


class Model:
    def __init__(self, inputs=None, outputs=None, name=None):
        self.inputs = inputs if isinstance(inputs, list) else [inputs]
        self.outputs = outputs if isinstance(outputs, list) else [outputs]
        self.input = self.inputs[0] if len(self.inputs) == 1 else self.inputs          # Keras: a list when there are several
        self.output = self.outputs[0] if len(self.outputs) == 1 else self.outputs
        self.name = name

    def get_layer(self, name):
        return LAYERS[name]


def install():
    KL.Conv3D, KL.BatchNormalization, KL.MaxPooling3D, KL.UpSampling3D = Conv3D, BatchNormalization, MaxPooling3D, UpSampling3D
    KL.Activation, KL.Input, KL.concatenate = Activation, Input, concatenate
    sys.modules['keras.models'].Model = Model


def main():
    install()
    sys.path.insert(0, '/root/reference')
    from ext.neuron import models as nrn_models
    out = {}
    # the configuration SynthSR.training() builds (training.py:330-341) at toy size: 3 levels, 4 features (spatial dims
    # divisible by 4, as the reference requires through output_div_by_n)
    shape = [12, 8, 16]
    image = RNG.normal(size=(2, *shape, 2)).astype(np.float32)
    FEED.append(KTensor(image))
    model = nrn_models.unet(nb_features=4, input_shape=[*shape, 2], nb_levels=3, conv_size=3, nb_labels=1, feat_mult=2,
                            nb_conv_per_level=2, conv_dropout=0, final_pred_activation='linear', batch_norm=-1,
                            activation='elu', input_model=None)
    out['image'] = image
    out['prediction'] = np.asarray(model.output, dtype=np.float64)
    out['layer_order'] = np.array(list(LAYERS))
    for n in ('unet_conv_downarm_0_1', 'unet_bn_down_0', 'unet_maxpool_0', 'unet_up_3', 'unet_merge_3', 'unet_bn_up_1',
              'unet_likelihood'):
        out['act/' + n] = np.asarray(LAYERS[n].output, dtype=np.float64)
    for k, v in WEIGHTS.items():
        out['w/' + k] = v
    print(len(LAYERS), 'layers;', len(WEIGHTS), 'weights; prediction', out['prediction'].shape)
    print(list(LAYERS))
    np.savez_compressed(os.path.join(HERE, 'reference_unet.npz'), **out)


if __name__ == '__main__':
    main()
