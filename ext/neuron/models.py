"""`ext.neuron.models.unet` of the reference (ext/neuron/models.py:26-145) on the B200 engine.

Returns a `UnetModel` that plays the role of the Keras `Model` for the training path: `.predict(image)`,
`.get_weights()/.set_weights()` by Keras layer name, `.save_weights()/.load_weights()` (Keras .h5 through the pure-Python
HDF5 reader/writer synthsr_b200/h5lite.py -- h5py is not installed -- or .npz).  Auto-encoder variants (`ae`, `single_ae`, `add_prior`) are not part
of SynthSR's path and are not provided."""
import numpy as np


class UnetModel:
    def __init__(self, net, input_model=None, name='unet'):
        self.net, self.input_model, self.name = net, input_model, name
        self.inputs = input_model.inputs if input_model is not None else ['%s_input' % name]
        self.output_shape = [None] + net.dims + [net.nb_labels]

    @property
    def layer_names(self):
        return [n for n, *_ in self.net.specs]

    def predict(self, image):
        """image: [B,X,Y,Z,C] numpy -> prediction numpy (inference mode: moving BN statistics)."""
        import torch
        x = torch.as_tensor(np.ascontiguousarray(image, dtype=np.float32)).cuda()
        return self.net.predict(x).cpu().numpy()

    def get_weights(self):
        return self.net.state_dict()

    def set_weights(self, sd):
        self.net.load_state_dict(sd, strict=False)

    def save_weights(self, path):
        """'.h5': Keras `save_weights` layout (readable by keras `load_weights(by_name=True)` and the reference's
        predict scripts); anything else: .npz with the same '<layer>/<weight>' keys."""
        if str(path).endswith('.h5'):
            from synthsr_b200 import h5lite
            from synthsr_b200.unet import keras_layer_order
            h5lite.save_keras_weights(path, self.net.state_dict(), keras_layer_order(self.net.L))
        else:
            np.savez(path, **self.net.state_dict())

    def load_weights(self, path, by_name=True):
        """Keras .h5 (`save_weights` or full `model.save` / ModelCheckpoint files, e.g. models/SynthSR_v10_210712.h5) or
        .npz.  by_name=True (the only mode the reference uses, training.py:362): layers are matched by name, layers
        absent from the file keep their values; a shape mismatch raises like Keras does."""
        if str(path).endswith('.h5'):
            from synthsr_b200 import h5lite
            sd, _ = h5lite.load_keras_weights(path)
        else:
            sd = {k: v for k, v in dict(np.load(path)).items() if not k.startswith('optimizer/')}
        for k, v in sd.items():
            if k in self.net.p and tuple(self.net.p[k].shape) != tuple(np.shape(v)):
                raise ValueError('Layer weight shape %s of %s not compatible with provided weight shape %s'
                                 % (tuple(self.net.p[k].shape), k, tuple(np.shape(v))))
        self.net.load_state_dict(sd, strict=not by_name)


def unet(nb_features, input_shape, nb_levels, conv_size, nb_labels, name='unet', prefix=None, feat_mult=1, pool_size=2,
         use_logp=True, padding='same', dilation_rate_mult=1, activation='elu', skip_n_concatenations=0,
         use_residuals=False, final_pred_activation='softmax', nb_conv_per_level=1, add_prior_layer=False,
         layer_nb_feats=None, conv_dropout=0, batch_norm=None, input_model=None, batchsize=1, conv_impl='tc', seed=None):
    """Same keyword names as the reference.  The engine implements the configuration SynthSR.training() uses
    (training.py:330-341): 'same' padding, ELU, batch_norm=-1, 2 convs per level, no residuals/dropout/dilation,
    linear final activation; anything else raises NotImplementedError instead of silently differing."""
    from synthsr_b200.unet import UNet3D
    unsupported = []
    if pool_size not in (2, (2, 2, 2), [2, 2, 2]): unsupported.append('pool_size')
    if padding != 'same': unsupported.append('padding')
    if dilation_rate_mult != 1: unsupported.append('dilation_rate_mult')
    if activation != 'elu': unsupported.append('activation')
    if skip_n_concatenations: unsupported.append('skip_n_concatenations')
    if use_residuals: unsupported.append('use_residuals')
    if final_pred_activation != 'linear': unsupported.append('final_pred_activation')
    if nb_conv_per_level != 2: unsupported.append('nb_conv_per_level')
    if add_prior_layer: unsupported.append('add_prior_layer')
    if layer_nb_feats is not None: unsupported.append('layer_nb_feats')
    if conv_dropout: unsupported.append('conv_dropout')
    if batch_norm != -1: unsupported.append('batch_norm')
    if unsupported:
        raise NotImplementedError('unet(): options outside the SynthSR training configuration: %s' % ', '.join(unsupported))
    if input_model is not None:
        batchsize = getattr(input_model, 'batchsize', batchsize)
    net = UNet3D(list(input_shape), nb_features=nb_features, nb_levels=nb_levels, conv_size=conv_size,
                 nb_labels=nb_labels, feat_mult=feat_mult, nb_conv_per_level=nb_conv_per_level, batchsize=batchsize,
                 conv_impl=conv_impl, seed=seed)
    return UnetModel(net, input_model, name)
