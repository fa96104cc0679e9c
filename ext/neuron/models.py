"""`ext.neuron.models.unet` of the reference (ext/neuron/models.py:26-145) on the B200 engine.

Returns a `UnetModel` that plays the role of the Keras `Model` for the training path: `.predict(image)`,
`.get_weights()/.set_weights()` by Keras layer name, `.save_weights()/.load_weights()` (Keras .h5 through the pure-Python
HDF5 reader/writer synthsr_b200/h5lite.py -- h5py is not installed -- or .npz).  Auto-encoder variants (`ae`, `single_ae`, `add_prior`) are not part
of SynthSR's path and are not provided."""
import numpy as np


class DynamicUnetModel:
    """`unet(input_shape=[None, None, None, C], ...)` of the inference scripts (scripts/predict_command_line.py:66-77):
    the spatial shape is only known per scan, so the engine (activation buffers, TMA descriptors) is built -- and cached --
    per input shape at predict() time; the weights live here by Keras layer name."""

    def __init__(self, cin, kwargs, name='unet'):
        self.cin, self.kwargs, self.name = int(cin), kwargs, name
        self.inputs = ['%s_input' % name]
        self.output_shape = [None, None, None, None, kwargs['nb_labels']]
        self._sd = None
        self._nets = {}

    @property
    def layer_names(self):
        from synthsr_b200.unet import layer_specs
        k = self.kwargs
        return [n for n, *_ in layer_specs(self.cin, k['nb_features'], k['nb_levels'], k['feat_mult'],
                                           k['nb_conv_per_level'], k['nb_labels'])]

    def _net(self, dims, batch):
        from synthsr_b200.unet import UNet3D
        key = (tuple(dims), batch)
        if key not in self._nets:
            self._nets.clear()                                # one shape resident at a time (full-size scans are large)
            net = UNet3D(list(dims) + [self.cin], batchsize=batch, seed=0, **self.kwargs)
            if self._sd is not None:
                net.load_state_dict(self._sd, strict=False)
            self._nets[key] = net
        return self._nets[key]

    def predict(self, image):
        """image [B,X,Y,Z,C] numpy (X, Y, Z multiples of 2**(nb_levels-1)) -> prediction, inference-mode BatchNorm."""
        import torch
        image = np.ascontiguousarray(image, dtype=np.float32)
        assert image.ndim == 5 and image.shape[-1] == self.cin, image.shape
        net = self._net(image.shape[1:4], image.shape[0])
        return net.predict(torch.as_tensor(image).cuda()).cpu().numpy()

    def get_weights(self):
        return dict(self._sd or {})

    def set_weights(self, sd):
        self._sd = dict(self._sd or {}, **{k: np.asarray(v, dtype=np.float32) for k, v in sd.items()})
        for net in self._nets.values():
            net.load_state_dict(self._sd, strict=False)

    def load_weights(self, path, by_name=True):
        if str(path).endswith('.h5'):
            from synthsr_b200 import h5lite
            sd, _ = h5lite.load_keras_weights(path)
        else:
            sd = {k: v for k, v in dict(np.load(path)).items() if not k.startswith('optimizer/')}
        from synthsr_b200.unet import layer_specs
        k = self.kwargs
        for name, kind, ci, co in layer_specs(self.cin, k['nb_features'], k['nb_levels'], k['feat_mult'],
                                              k['nb_conv_per_level'], k['nb_labels']):
            if kind != 'bn' and name + '/kernel' in sd:
                ks = k['conv_size'] if kind == 'conv' else 1
                if tuple(sd[name + '/kernel'].shape) != (ks, ks, ks, ci, co):
                    raise ValueError('Layer weight shape %s of %s not compatible with provided weight shape %s'
                                     % ((ks, ks, ks, ci, co), name, tuple(sd[name + '/kernel'].shape)))
        self.set_weights(sd)

    def save_weights(self, path):
        from synthsr_b200 import h5lite
        from synthsr_b200.unet import keras_layer_order
        if str(path).endswith('.h5'):
            h5lite.save_keras_weights(path, self._sd or {}, keras_layer_order(self.kwargs['nb_levels']))
        else:
            np.savez(path, **(self._sd or {}))


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
         layer_nb_feats=None, conv_dropout=0, batch_norm=None, input_model=None, batchsize=1, conv_impl='tc3', seed=None):
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
    if (prefix if prefix is not None else name) != 'unet': unsupported.append('name/prefix (layer names are unet_*)')
    if unsupported:
        raise NotImplementedError('unet(): options outside the SynthSR training configuration: %s' % ', '.join(unsupported))
    if input_model is not None:
        batchsize = getattr(input_model, 'batchsize', batchsize)
    if any(d is None for d in list(input_shape)[:3]):         # inference scripts: spatial shape known per scan only
        return DynamicUnetModel(input_shape[3], dict(nb_features=nb_features, nb_levels=nb_levels, conv_size=conv_size,
                                                     nb_labels=nb_labels, feat_mult=feat_mult,
                                                     nb_conv_per_level=nb_conv_per_level, conv_impl=conv_impl), name)
    net = UNet3D(list(input_shape), nb_features=nb_features, nb_levels=nb_levels, conv_size=conv_size,
                 nb_labels=nb_labels, feat_mult=feat_mult, nb_conv_per_level=nb_conv_per_level, batchsize=batchsize,
                 conv_impl=conv_impl, seed=seed)
    return UnetModel(net, input_model, name)
