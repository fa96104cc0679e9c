"""Golden vectors for ONE STEP OF THE WHOLE TRAINING GRAPH as the reference's own SynthSR/training.training() builds it:
get_list_labels -> BrainGenerator (-> labels_to_image_model on the tf shim) -> ext.neuron.models.unet(input_model=...) on the
functional Keras stand-in -> metrics_model, all unmodified; only train_model (Keras compile / fit_generator) is replaced by a
function that captures the model whose output, in this eager harness, IS the loss of the fed batch.  The inputs fed to
KL.Input are what the reference's own build_model_inputs yields for the same label maps and priors.

Stored: the batch (labels, means, stds), every tf.random draw (as the `draws` dict), the U-Net weights by Keras name, the
generator outputs, the prediction and the loss -- plus what training() derived on the way (padding margin, output shape,
residual channel list) for the test of the product's argument handling.

Cases:
  plain     training() defaults on 1 channel, output_shape 16, l1, no loss cropping
  residual  2 input channels + reliability maps, work_with_residual_channel=[0], loss_cropping 8, l2
            (the reference's `2 * list` at training.py:270-271 REPEATS the list: [0] -> [0, 0])

Writes tests/golden/reference_training.npz.   (build container only: needs /root/reference)"""
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_model_goldens as MG  # noqa: E402  (tf shim with logged draws, KL.Input feed)
import make_reference_unet_goldens as UG  # noqa: E402   (functional Keras stand-in)

shim, T, f32 = MG.shim, MG.T, np.float32
K = sys.modules['keras.backend']
KL = sys.modules['keras.layers']
UG.install()


def Input(shape=None, name=None, dtype=None):              # fed arrays, registered under the input layer's name
    x = MG.FEED.pop(0)
    layer = types.SimpleNamespace(name=name, output=x)
    UG.LAYERS[name] = layer
    return x


class Lambda:                                               # KL.Lambda(fn, name=...)(x): evaluated, registered by name
    def __init__(self, fn, name=None, **kw):
        self.fn, self.name = fn, name

    def __call__(self, x):
        self.output = self.fn(x)
        if self.name is not None:
            UG.LAYERS[self.name] = self
        return self.output


class _Merge:
    def __init__(self, name=None, **kw):
        self.name = name

    def __call__(self, xs):
        self.output = T(self.op(np.asarray(xs[0], dtype=np.float64), np.asarray(xs[1], dtype=np.float64)))
        if self.name is not None:
            UG.LAYERS[self.name] = self
        return self.output


class Add(_Merge):
    op = staticmethod(lambda a, b: a + b)                   # NumPy broadcasting == what Keras does for [.., 2] + [.., 1]? no:
    # Keras' Add requires equal shapes except for broadcastable 1-dims, which is the case the reference produces here


class Subtract(_Merge):
    op = staticmethod(lambda a, b: a - b)


KL.Input, KL.Lambda, KL.Add, KL.Subtract = Input, Lambda, Add, Subtract
K.mean = lambda x, axis=None: T(np.mean(np.asarray(x, dtype=np.float64), axis=axis))
K.abs = lambda x: T(np.abs(np.asarray(x)))
sys.modules['keras'].models = sys.modules['keras.models']

import SynthSR.training as RT  # noqa: E402  (the reference's)
import SynthSR.brain_generator as RBG  # noqa: E402
from SynthSR.model_inputs import build_model_inputs  # noqa: E402

CAPTURED = {}


def capture_train_model(model, generator, lr, lr_decay, epochs, steps, model_dir, checkpoint=None):
    CAPTURED.update(model=model, lr=lr, lr_decay=lr_decay, epochs=epochs, steps=steps)


RT.train_model = capture_train_model
RT.models = sys.modules['keras.models']          # `from keras import models` ran when the package was first imported
_mm = RT.metrics_model


def capture_metrics_model(**kw):
    CAPTURED['metrics_kwargs'] = {k: v for k, v in kw.items() if k != 'input_model'}
    return _mm(**kw)


RT.metrics_model = capture_metrics_model
_l2i = RBG.labels_to_image_model


def capture_labels_to_image_model(**kw):                    # the exact keyword arguments BrainGenerator hands to the graph builder
    CAPTURED['l2i_kwargs'] = dict(kw)
    return _l2i(**kw)


RBG.labels_to_image_model = capture_labels_to_image_model
_unet = RT.nrn_models.unet


def capture_unet(**kw):
    CAPTURED['unet_kwargs'] = {k: v for k, v in kw.items() if k != 'input_model'}
    return _unet(**kw)


RT.nrn_models = types.SimpleNamespace(unet=capture_unet)


def run_case(name, seed, labels_shape, n_maps, kw):
    rng = np.random.default_rng(seed)
    MG.RNG[0] = rng
    del MG.LOG[:], MG.FEED[:], MG.FORCED[:]
    UG.LAYERS.clear()
    UG.WEIGHTS.clear()
    CAPTURED.clear()
    tmp = tempfile.mkdtemp()
    labels_dir = os.path.join(tmp, 'labels')
    os.makedirs(labels_dir)
    g = np.stack(np.meshgrid(*[np.linspace(0, 3, s) for s in labels_shape], indexing='ij'), -1)
    maps = []
    for i in range(n_maps):
        ph = rng.uniform(0, 6, size=3)
        fld = np.sin(g[..., 0] * 2 + ph[0]) + np.cos(g[..., 1] * 3 + ph[1]) + np.sin(g[..., 2] * 2.5 + ph[2])
        m = MG.GEN[np.clip(((fld + 3) / 6 * len(MG.GEN)).astype(int), 0, len(MG.GEN) - 1)].astype(np.int32)
        m[0, 0, :len(MG.GEN)] = MG.GEN                        # every label present in every map
        np.savez(os.path.join(labels_dir, 'map%d.npz' % i), vol_data=m)
        maps.append(m)
    n_ch = len(kw['input_channels'])
    K_ = len(MG.GEN)
    pm = np.concatenate([np.stack([rng.uniform(40, 200, size=K_), rng.uniform(2, 20, size=K_)]) for _ in range(n_ch)])
    ps = np.concatenate([np.stack([rng.uniform(5, 20, size=K_), rng.uniform(1, 4, size=K_)]) for _ in range(n_ch)])
    gen_labels_path = os.path.join(tmp, 'generation_labels.npy')
    np.save(gen_labels_path, MG.GEN)
    # the batch the reference's own input sampler yields first
    from ext.lab2im import utils
    paths = utils.list_images_in_folder(labels_dir)
    np.random.seed(seed)
    inputs = next(build_model_inputs(paths, K_, pm, ps, 'normal', batchsize=1, n_channels=n_ch))
    inputs = [np.asarray(a) for a in inputs]
    inputs[0] = inputs[0].astype(np.int32)
    inputs[1], inputs[2] = inputs[1].astype(f32), inputs[2].astype(f32)
    MG.FEED.extend(T(a) for a in inputs)
    MG.FORCED.extend(kw.pop('forced'))
    shim.base.GRAPH_BATCH[0] = 1
    RT.training(labels_dir=labels_dir, model_dir=os.path.join(tmp, 'models'), prior_means=pm, prior_stds=ps,
                path_generation_labels=gen_labels_path, FS_sort=False, **kw)
    shim.base.GRAPH_BATCH[0] = None
    assert not MG.FEED and not MG.FORCED
    log = [(k, np.asarray(v), a, b) for k, v, a, b in MG.LOG]
    image = np.asarray(UG.LAYERS['image_out'].output)
    target = np.asarray(UG.LAYERS['regression_target'].output)
    pmg = CAPTURED['metrics_kwargs']
    lk = CAPTURED['l2i_kwargs']
    cfg = {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in lk.items()
           if k not in ('labels_shape', 'generation_labels', 'n_neutral_labels')}
    assert list(lk['labels_shape']) == list(labels_shape) and list(lk['generation_labels']) == list(MG.GEN)
    pm_ = lk.get('padding_margin')
    pm_ = [0, 0, 0] if pm_ is None else (list(np.ravel(pm_)) * 3)[:3]
    grid = [int(s + 2 * m) for s, m in zip(labels_shape, pm_)]
    crop = [v for k, v, _, _ in log if k == 'normal' and v.ndim == 5][1 if lk['nonlin_std'] > 0 else 0].shape[1:4]
    draws = MG.to_draws(cfg, log, 1, list(crop) != grid)
    out = {'%s_in%d' % (name, i): a for i, a in enumerate(inputs)}
    out.update({'%s_draw_%s' % (name, k): np.asarray(v) for k, v in draws.items() if v is not None})
    out.update({'%s_w/%s' % (name, k): v for k, v in UG.WEIGHTS.items()})
    out['%s_image' % name], out['%s_target' % name] = image, target
    out['%s_prediction' % name] = np.asarray(UG.LAYERS['unet_prediction'].output, dtype=np.float64)
    out['%s_loss' % name] = np.asarray(CAPTURED['model'].outputs, dtype=np.float64).reshape(())
    meta = dict(cfg=cfg, n_neutral_labels=int(lk['n_neutral_labels']), labels_shape=list(labels_shape), grid_shape=grid,
                crop_shape=list(crop), image_shape=list(image.shape), metrics_kwargs=pmg, unet_kwargs=CAPTURED['unet_kwargs'],
                lr=CAPTURED['lr'], lr_decay=CAPTURED['lr_decay'], epochs=CAPTURED['epochs'], steps=CAPTURED['steps'],
                layers=[n for n in UG.LAYERS], training_kwargs={k: v for k, v in kw.items()})
    print(name, 'image', image.shape, 'target', target.shape, 'loss', float(out['%s_loss' % name]), pmg)
    return out, meta


CASES = {
    'plain': dict(seed=71, labels_shape=(20, 24, 20), n_maps=2, kw=dict(
        forced=[.3], randomise_res=False, input_channels=[True], output_channel=0, output_shape=16, n_levels=2, unet_feat_count=4,
        regression_metric='l1', loss_cropping=None, lr=2e-4, lr_decay=1e-6)),
    'residual': dict(seed=72, labels_shape=(18, 18, 22), n_maps=2, kw=dict(
        forced=[.3, .4], randomise_res=False, input_channels=[True, True], output_channel=[1], output_shape=16, n_levels=2, unet_feat_count=4,
        build_reliability_maps=True, work_with_residual_channel=[0], regression_metric='l2', loss_cropping=8)),
}

if __name__ == '__main__':
    out, meta = {}, {}
    for name, c in CASES.items():
        o, m = run_case(name, c['seed'], c['labels_shape'], c['n_maps'], dict(c['kw']))
        out.update(o)
        meta[name] = m
    out['generation_labels'] = MG.GEN
    out['meta_json'] = np.frombuffer(json.dumps(meta, default=lambda o: o.tolist() if hasattr(o, 'tolist') else str(o)).encode(),
                                     dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, 'reference_training.npz'), **out)
