"""One step of the WHOLE training graph as the reference's own SynthSR/training.training() builds it (get_list_labels ->
BrainGenerator -> labels_to_image_model -> ext.neuron.models.unet(input_model=...) -> metrics_model; executed by
tests/golden/make_reference_training_goldens.py with only Keras' compile / fit replaced) against
  * the oracle chain labels_to_image -> unet.forward -> loss_fn on the same batch, draws and weights;
  * the product's SynthSR.training.training() argument handling (what it hands to the generator plan and to the engine)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import generator as OG
from oracle import unet as OU

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, 'golden', 'reference_training.npz'))
META = json.loads(bytes(G['meta_json']).decode())


def _case(name):
    m = META[name]
    cfg = dict(m['cfg'])
    cfg['generation_labels'] = G['generation_labels']
    cfg['n_neutral_labels'] = m['n_neutral_labels']
    for k in ('data_res', 'thickness', 'aff'):
        if cfg.get(k) is not None:
            cfg[k] = np.array(cfg[k])
    inputs = [G['%s_in%d' % (name, i)] for i in range(3)]
    pre = '%s_draw_' % name
    draws = {k[len(pre):]: G[k] for k in G.files if k.startswith(pre)}
    for k in ('aff_rotation', 'aff_shearing', 'aff_scaling', 'aff_translation'):
        draws.setdefault(k, None)
    for k in list(draws):
        if k.startswith('bias_apply'):
            draws[k] = bool(draws[k])
    pre = '%s_w/' % name
    weights = {k[len(pre):]: G[k] for k in G.files if k.startswith(pre)}
    return m, cfg, inputs, draws, weights


@pytest.mark.parametrize('name', sorted(META))
def test_oracle_chain_reproduces_the_reference_training_graph(name):
    m, cfg, inputs, draws, weights = _case(name)
    image, target = OG.labels_to_image(cfg, inputs, draws)
    np.testing.assert_allclose(image, G[name + '_image'], rtol=0, atol=2e-6)
    np.testing.assert_allclose(target, G[name + '_target'], rtol=0, atol=2e-6)
    uk = m['unet_kwargs']
    assert uk['input_shape'] == list(image.shape[1:]) and uk['batch_norm'] == -1 and uk['final_pred_activation'] == 'linear'
    kw = dict(nb_features=uk['nb_features'], nb_levels=uk['nb_levels'], feat_mult=uk['feat_mult'],
              nb_conv_per_level=uk['nb_conv_per_level'], nb_labels=uk['nb_labels'])
    params = OU.init_params(0, image.shape[-1], dtype=torch.float64, **kw)
    assert {k for k in params if not k.endswith(('moving_mean', 'moving_variance'))} == set(weights)
    for k, v in weights.items():
        params[k] = torch.from_numpy(v.astype(np.float64))
    # from the reference's generator outputs, so that the U-Net / loss comparison is not blurred by the generator's few ulp
    img_t, tgt_t = torch.from_numpy(G[name + '_image'].astype(np.float64)), torch.from_numpy(G[name + '_target'].astype(np.float64))
    pred = OU.forward(params, img_t, training=True, nb_levels=uk['nb_levels'])
    np.testing.assert_allclose(pred.numpy(), G[name + '_prediction'], rtol=0, atol=1e-9)
    mk = m['metrics_kwargs']
    loss = OU.loss_fn(pred, img_t, tgt_t, metric=mk['metrics'], work_with_residual_channel=mk['work_with_residual_channel'],
                      loss_cropping=mk['loss_cropping'])
    np.testing.assert_allclose(float(loss), float(G[name + '_loss']), rtol=1e-10)
    # end to end through the oracle's own generator output as well
    pred2 = OU.forward(params, torch.from_numpy(image.astype(np.float64)), training=True, nb_levels=uk['nb_levels'])
    loss2 = OU.loss_fn(pred2, torch.from_numpy(image.astype(np.float64)), torch.from_numpy(target.astype(np.float64)),
                       metric=mk['metrics'], work_with_residual_channel=mk['work_with_residual_channel'],
                       loss_cropping=mk['loss_cropping'])
    np.testing.assert_allclose(float(loss2), float(G[name + '_loss']), rtol=1e-4)


def test_reference_repeats_the_residual_list_and_the_step_is_that_of_the_undoubled_index():
    """training.py:270-271: `2 * work_with_residual_channel` -> [0, 0]; loss and gradient over two identical copies equal those
    of the single un-doubled channel, which is what the product hands to its engine."""
    m, cfg, inputs, draws, weights = _case('residual')
    assert m['metrics_kwargs']['work_with_residual_channel'] == [0, 0]
    img = torch.from_numpy(G['residual_image'].astype(np.float64))
    tgt = torch.from_numpy(G['residual_target'].astype(np.float64))
    pred = torch.from_numpy(G['residual_prediction']).clone().requires_grad_(True)
    l2 = OU.loss_fn(pred, img, tgt, metric='l2', work_with_residual_channel=[0, 0], loss_cropping=8)
    g2, = torch.autograd.grad(l2, pred)
    pred1 = pred.detach().clone().requires_grad_(True)
    l1 = OU.loss_fn(pred1, img, tgt, metric='l2', work_with_residual_channel=[0], loss_cropping=8)
    g1, = torch.autograd.grad(l1, pred1)
    np.testing.assert_allclose(float(l1.detach()), float(l2.detach()), rtol=1e-12)
    np.testing.assert_allclose(g1.numpy(), g2.numpy(), rtol=0, atol=1e-15)


class _CapturedEngine:
    last = None

    def __init__(self, plan, **kw):
        self.plan, self.kw = plan, kw
        self.net = None
        _CapturedEngine.last = self


@pytest.mark.parametrize('name', sorted(META))
def test_product_training_derives_what_the_reference_training_derives(name, tmp_path, monkeypatch):
    """SynthSR.training.training() of the product, with the engine replaced by a recorder: generator plan (padding, crop and
    output shapes, channel counts, every augmentation hyper-parameter) and engine arguments (U-Net size, loss, learning rate)
    against what the reference's training() handed to labels_to_image_model / unet / metrics_model / train_model."""
    import SynthSR.training as PT
    import synthsr_b200.trainer as TR
    m = META[name]
    labels_dir = tmp_path / 'labels'
    labels_dir.mkdir()
    rng = np.random.default_rng(3)
    gen = G['generation_labels']
    for i in range(2):
        lab = gen[rng.integers(0, len(gen), size=m['labels_shape'])].astype(np.int32)
        lab[0, 0, :len(gen)] = gen
        np.savez(str(labels_dir / ('map%d.npz' % i)), vol_data=lab)
    np.save(str(tmp_path / 'gen.npy'), gen)
    monkeypatch.setattr(TR, 'TrainingEngine', _CapturedEngine)
    monkeypatch.setattr(PT, 'train_model', lambda *a, **k: None)
    monkeypatch.setattr(PT, 'metrics_model', lambda *a, **k: None)
    monkeypatch.setattr(PT.nrn_models, 'UnetModel', lambda *a, **k: None)
    kw = dict(m['training_kwargs'])
    K_ = len(gen)
    n_ch = len(kw['input_channels'])
    pm = np.tile(np.array([[100.] * K_, [10.] * K_]), (n_ch, 1))
    PT.training(str(labels_dir), str(tmp_path / 'models'), pm, pm, str(tmp_path / 'gen.npy'), FS_sort=False, **kw)
    eng = _CapturedEngine.last
    p, cfg = eng.plan, m['cfg']
    assert p.grid_shape == m['grid_shape'] and p.crop_shape == m['crop_shape']
    assert p.image_shape == m['image_shape'][1:]
    assert p.n_neutral_labels == m['n_neutral_labels']
    for key in ('scaling_bounds', 'rotation_bounds', 'shearing_bounds', 'translation_bounds', 'nonlin_std',
                'nonlin_shape_factor', 'blur_range', 'bias_field_std', 'build_reliability_maps', 'flipping'):
        assert getattr(p, key) == cfg[key], key
    assert p.bias_small_shape == OG.get_resample_shape(m['crop_shape'], cfg['bias_shape_factor'])
    assert [bool(v) for v in p.downsample] == [bool(v) for v in np.ravel(cfg['downsample'])] or \
        [bool(v) for v in p.downsample] == [bool(cfg['downsample'])] * p.n_channels
    assert p.sim_reg == ([bool(cfg['simulate_registration_error'])] * p.n_channels)
    assert p.randomise_res == [bool(cfg['randomise_res'])] * p.n_channels
    uk, mk = m['unet_kwargs'], m['metrics_kwargs']
    assert eng.kw['nb_features'] == uk['nb_features'] and eng.kw['nb_levels'] == uk['nb_levels']
    assert eng.kw['conv_size'] == uk['conv_size'] and eng.kw['feat_mult'] == uk['feat_mult']
    assert eng.kw['nb_conv_per_level'] == uk['nb_conv_per_level'] and eng.kw['nb_labels'] == uk['nb_labels']
    assert eng.kw['lr'] == m['lr'] and eng.kw['lr_decay'] == m['lr_decay']
    assert eng.kw['metric'] == mk['metrics'] and eng.kw['loss_cropping'] == mk['loss_cropping']
    ref_res = mk['work_with_residual_channel']
    assert eng.kw['work_with_residual_channel'] == (None if ref_res is None else sorted(set(ref_res)))


def test_residual_on_several_outputs_with_reliability_maps_fails_like_keras(tmp_path):
    from SynthSR.training import training
    with pytest.raises(ValueError, match='could not be broadcast'):
        training('x', str(tmp_path), None, None, None, input_channels=[True, True], output_channel=[0, 1],
                 work_with_residual_channel=[0, 1], build_reliability_maps=True)
