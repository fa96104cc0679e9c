"""`ext.neuron.models.unet` of the reference (ext/neuron/models.py:26-145) on the B200 engine.

Returns a `UnetModel` that plays the role of the Keras `Model` for the training path: `.predict(image)`,
`.get_weights()/.set_weights()` by Keras layer name, `.save_weights()/.load_weights()` (.npz; an .h5 writer needs
h5py, which this image lacks -- SURVEY.md 8f).  Auto-encoder variants (`ae`, `single_ae`, `add_prior`) are not part
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
        np.savez(path, **self.net.state_dict())

    def load_weights(self, path, by_name=True):
        if path.endswith('.h5'):
            raise NotImplementedError('reading Keras .h5 needs an HDF5 reader (h5py is not installed); convert the '
                                      'file to .npz with the Keras layer names (kernel/bias/gamma/beta/moving_*)')
        sd = dict(np.load(path))
        self.net.load_state_dict({k: v for k, v in sd.items() if not k.startswith('optimizer/')}, strict=not by_name)


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
