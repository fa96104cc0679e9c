"""CPU suite: the product generator's HOST orchestration (synthsr_b200/generator.py: which entry point runs on which buffer,
with which shapes, strides and parameter blocks) executed against tests/host_emulator.py -- a NumPy stand-in for the
generator entry points of the C ABI -- and compared with
  (1) the reference's own graph outputs (tests/golden/reference_model.npz, made by executing
      SynthSR/labels_to_image_model.labels_to_image_model on the tf shim), same inputs and same draws;
  (2) the oracle's end-to-end graph on the configurations tests/test_generator_gpu.py runs on the B200.
The CUDA kernels are NOT exercised here (tests/test_generator_gpu.py does that through the real library); what this
file pins is everything around them, e.g. that a channel which is both input and target continues on the output grid."""
import numpy as np
import pytest
import torch

from helpers import GEN_LABELS, gmm_params, phantom_labels
from host_emulator import HostEmulator
from oracle import generator as OG
from test_reference_model_goldens import ATOL, META, _case

f32 = np.float32


@pytest.fixture
def emu(monkeypatch):
    import synthsr_b200.generator as G
    e = HostEmulator()
    monkeypatch.setattr(G, 'lib', e)
    monkeypatch.setattr(G, 'stream_ptr', lambda: 0)
    return e


def _plan(cfg, labels_shape, label_list):
    from synthsr_b200.generator import GeneratorPlan
    skip = ('input_channels', 'output_channel', 'n_neutral_labels', 'atlas_res', 'target_res', 'generation_labels')
    return GeneratorPlan(labels_shape, cfg.get('input_channels', True), cfg.get('output_channel', 0), label_list,
                         cfg.get('n_neutral_labels'), cfg.get('atlas_res', 1.), cfg.get('target_res'),
                         **{k: v for k, v in cfg.items() if k not in skip})


def _run(plan, inputs, draws, batch):
    from synthsr_b200.generator import SynthGenerator
    gen = SynthGenerator(plan, batchsize=batch, device='cpu')
    labels = torch.from_numpy(np.ascontiguousarray(inputs[0][..., 0].astype(np.int32)))
    real = torch.from_numpy(np.ascontiguousarray(inputs[3][..., 0])) if len(inputs) > 3 else None
    keep = {}
    image, target = gen.run(labels, inputs[1], inputs[2], draws, real_image=real, keep=keep)
    return image.numpy().copy(), target.numpy().copy(), keep


def test_product_refuses_a_cpu_device_without_the_emulator():
    from synthsr_b200.generator import SynthGenerator
    plan = _plan(dict(), [16, 16, 16], GEN_LABELS)
    with pytest.raises(RuntimeError, match='CUDA device'):
        SynthGenerator(plan, batchsize=1, device='cpu')


@pytest.mark.parametrize('tag', sorted(META))
def test_orchestration_reproduces_the_reference_graph(emu, tag):
    """product host code + emulated entry points == the reference's own labels_to_image_model outputs."""
    cfg, inputs, draws, ref_image, ref_target = _case(tag)
    plan = _plan(cfg, META[tag]['labels_shape'], cfg['generation_labels'])
    image, target, _ = _run(plan, inputs, draws, META[tag]['batch'])
    assert image.shape == ref_image.shape and target.shape == ref_target.shape
    np.testing.assert_allclose(image, ref_image, rtol=0, atol=ATOL)
    np.testing.assert_allclose(target, ref_target, rtol=0, atol=ATOL)


def test_rebound_target_channel_runs_its_chain_on_the_output_grid(emu):
    """case E: channel 1 is input and target at target_res 1.5 -> registration warp, acquisition blur, nearest down-sampling
    and the reliability map of that channel are all launched with the OUTPUT grid's dimensions (labels_to_image_model.py
    :193-195 rebinds `channel`); channel 0 keeps the crop grid."""
    cfg, inputs, draws, _, _ = _case('E')
    plan = _plan(cfg, META['E']['labels_shape'], cfg['generation_labels'])
    assert plan.chan_grid == [plan.crop_shape, plan.output_shape] and plan.crop_shape != plan.output_shape
    _run(plan, inputs, draws, 1)
    warps = [a for n, a in emu.calls if n == 'ssr_warp_linear']
    assert [w[0] for w in warps] == [tuple(plan.output_shape)] * 3        # T, then Terr.Tinv on the channel and on its map
    downs = [a for n, a in emu.calls if n == 'ssr_resize' and a[3] == 1]  # nearest
    assert [d[0] for d in downs] == [tuple(plan.crop_shape), tuple(plan.output_shape)]
    assert [list(d[1]) for d in downs] == plan.down_shape


CONFIGS = [   # the configurations of tests/test_generator_gpu.py, at sizes the NumPy emulation finishes in seconds
    ('crop', dict(output_shape=16, translation_bounds=5, aff=np.eye(4)), [24, 24, 20], 1, False),
    ('batch2', dict(scaling_bounds=.15, rotation_bounds=15, shearing_bounds=.02, translation_bounds=5, nonlin_std=4.),
     [16, 20, 16], 2, False),
    ('multichannel', dict(input_channels=[False, True, True], output_channel=0, data_res=np.array([[1., 1., 3.], [1., 1., 4.]]),
                          thickness=np.array([[1., 1., 2.], [1., 1., 4.]]), downsample=True, build_reliability_maps=True,
                          output_shape=16), [20, 20, 24], 1, False),
    ('randomise_res', dict(input_channels=[True, True], output_channel=0, randomise_res=True, build_reliability_maps=True,
                           simulate_registration_error=True, output_shape=16), [20, 20, 20], 1, False),
    ('target_res', dict(target_res=2., padding_margin=4, nonlin_std=2.), [16, 24, 16], 1, False),
    ('target_res_finer', dict(target_res=.5, build_reliability_maps=True, data_res=np.array([[1., 2., 1.]]),
                              thickness=np.array([[1., 2., 1.]]), downsample=True), [12, 10, 12], 1, False),
    ('identity', dict(scaling_bounds=False, rotation_bounds=False, shearing_bounds=False, translation_bounds=False,
                      nonlin_std=0., flipping=False), [16, 16, 16], 1, False),
    ('real_image', dict(output_channel=None, output_shape=16), [20, 20, 18], 1, True),
]


@pytest.mark.parametrize('name,cfg,shape,batch,real', CONFIGS, ids=[c[0] for c in CONFIGS])
def test_orchestration_matches_oracle(emu, name, cfg, shape, batch, real):
    from synthsr_b200.draws import sample_draws
    rng = np.random.default_rng(5)
    labs = np.stack([phantom_labels(shape, GEN_LABELS, seed=3 + b) for b in range(batch)])
    plan = _plan(cfg, shape, GEN_LABELS)
    means, stds = gmm_params(rng, len(GEN_LABELS), plan.n_channels, batch)
    draws = sample_draws(rng, plan, batch, gmm_noise=True)
    inputs = [labs[..., None], means, stds]
    if real:
        inputs.append(rng.uniform(0, 200, size=(batch, *shape, 1)).astype(f32))
    image, target, keep = _run(plan, inputs, draws, batch)
    ocfg = dict(cfg, generation_labels=GEN_LABELS)
    o_image, o_target, inter = OG.labels_to_image(ocfg, inputs, draws, return_intermediates=True)
    np.testing.assert_array_equal(keep['labels'][0].numpy(), inter['labels'])
    np.testing.assert_allclose(image, o_image, rtol=0, atol=ATOL)
    np.testing.assert_allclose(target, o_target, rtol=0, atol=ATOL)


@pytest.fixture
def small_dataset(tmp_path):
    from ext.lab2im import utils
    from synthsr_b200.synthetic import GEN_CLASSES, GEN_LABELS as GL, phantom_labels as pl, synthetic_priors
    labels_dir = tmp_path / 'labels'
    labels_dir.mkdir()
    aff = np.array([[1., 0, 0, 40], [0, 1., 0, 16], [0, 0, 1., 63], [0, 0, 0, 1]])
    for i in range(2):
        utils.save_volume(pl([24, 28, 20], seed=i).astype(np.float32), aff, None, str(labels_dir / ('brain%d_labels.nii.gz' % i)))
    pm, ps = synthetic_priors(14, 1, seed=0)
    paths = {k: str(tmp_path / (k + '.npy')) for k in ('labels', 'classes', 'means', 'stds')}
    np.save(paths['labels'], GL); np.save(paths['classes'], GEN_CLASSES); np.save(paths['means'], pm); np.save(paths['stds'], ps)
    return str(labels_dir), paths


def test_brain_generator_generate_brain_on_the_emulated_entry_points(emu, small_dataset, monkeypatch):
    """the user-facing path of tutorials 1-6 -- BrainGenerator -> build_model_inputs -> labels_to_image_model.predict ->
    generate_brain (back to native space) -- end to end in the CPU suite (the GPU twin is tests/test_training_api_gpu.py)."""
    import functools
    import synthsr_b200.generator as G
    from SynthSR.brain_generator import BrainGenerator
    monkeypatch.setattr(G, 'SynthGenerator', functools.partial(G.SynthGenerator, device='cpu'))
    labels_dir, p = small_dataset
    gen = BrainGenerator(labels_dir, p['means'], p['stds'], 'normal', p['labels'], generation_classes=p['classes'],
                         output_shape=16, data_res=np.array([1., 1., 3.]), thickness=np.array([1., 1., 3.]),
                         downsample=True, build_reliability_maps=True)
    assert gen.labels_shape == [24, 28, 20] and gen.n_dims == 3 and gen.model_output_shape == [16, 16, 16, 2]
    image, target = gen.generate_brain()
    assert image.shape == (16, 16, 16, 2) and target.shape == (16, 16, 16)
    assert image.dtype == np.float32 and np.isfinite(image).all() and np.isfinite(target).all()
    assert 0. <= target.min() and target.max() <= 1. + 1e-6 and target.std() > 0.01
    rel = image[..., 1]
    assert rel.min() >= 0 and rel.max() <= 1 + 1e-6 and (rel < 0.99).any()      # interpolated slices are marked
    image2, _ = gen.generate_brain()
    assert not np.array_equal(image, image2)                                    # fresh augmentation every call
    names = [n for n, _ in emu.calls]
    assert names.count('ssr_deform_labels_nearest') == 2 and names.count('ssr_gmm_bias_minmax') == 2
