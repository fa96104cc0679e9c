"""End-to-end through the drop-in boundary on a GPU: SynthSR.training.training() and BrainGenerator.generate_brain()
on small NIfTI label maps written by the repo's own writer (nothing is read from /root/reference)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dataset(tmp_path_factory):
    from ext.lab2im import utils
    from synthsr_b200.synthetic import GEN_CLASSES, GEN_LABELS, phantom_labels, synthetic_priors
    root = tmp_path_factory.mktemp('data')
    labels_dir = root / 'labels'
    os.makedirs(labels_dir)
    aff = np.array([[1., 0, 0, 40], [0, 1., 0, 16], [0, 0, 1., 63], [0, 0, 0, 1]])     # like the reference fixtures
    for i in range(2):
        utils.save_volume(phantom_labels([44, 52, 40], seed=i).astype(np.float32), aff, None,
                          str(labels_dir / ('brain%d_labels.nii.gz' % i)))
    pm, ps = synthetic_priors(14, 1, seed=0)
    pm2, ps2 = synthetic_priors(14, 2, seed=1)
    paths = {k: str(root / (k + '.npy')) for k in ('labels', 'classes', 'means', 'stds', 'means2', 'stds2')}
    np.save(paths['labels'], GEN_LABELS); np.save(paths['classes'], GEN_CLASSES)
    np.save(paths['means'], pm); np.save(paths['stds'], ps); np.save(paths['means2'], pm2); np.save(paths['stds2'], ps2)
    return str(labels_dir), paths, root


def test_brain_generator_generate_brain(dataset):
    """tutorial-2 style: BrainGenerator(...).generate_brain() -> (image, target) in native space."""
    from SynthSR.brain_generator import BrainGenerator
    labels_dir, p, _ = dataset
    gen = BrainGenerator(labels_dir, p['means'], p['stds'], 'normal', p['labels'], generation_classes=p['classes'],
                         output_shape=32, data_res=np.array([1., 1., 3.]), thickness=np.array([1., 1., 3.]),
                         downsample=True, build_reliability_maps=True)
    assert gen.labels_shape == [44, 52, 40] and gen.n_dims == 3 and gen.model_output_shape == [32, 32, 32, 2]
    image, target = gen.generate_brain()
    assert image.shape == (32, 32, 32, 2) and target.shape == (32, 32, 32)
    assert image.dtype == np.float32 and np.isfinite(image).all() and np.isfinite(target).all()
    assert 0. <= target.min() and target.max() <= 1. + 1e-6 and target.std() > 0.01
    rel = image[..., 1]
    assert rel.min() >= 0 and rel.max() <= 1 + 1e-6 and (rel < 0.99).any()      # interpolated slices are marked
    image2, _ = gen.generate_brain()
    assert not np.array_equal(image, image2)                                    # fresh augmentation every call


def test_brain_generator_randomise_res(dataset):
    """tutorial-4 style (fine_tuning_with_adversary.py:64 uses it too): randomise_res=True -> random acquisition
    resolution per call, distance-to-acquired-voxel map as the reliability channel."""
    from SynthSR.brain_generator import BrainGenerator
    labels_dir, p, _ = dataset
    gen = BrainGenerator(labels_dir, p['means'], p['stds'], 'normal', p['labels'], generation_classes=p['classes'],
                         output_shape=32, randomise_res=True, build_reliability_maps=True)
    assert gen.model_output_shape == [32, 32, 32, 2]
    dmax = []
    for _ in range(4):
        image, target = gen.generate_brain()
        assert image.shape == (32, 32, 32, 2) and np.isfinite(image).all() and np.isfinite(target).all()
        dist = image[..., 1]
        assert dist.min() >= 0 and dist.max() <= np.sqrt(3 * 4.5 ** 2) + 1e-3     # at most half a 9 mm voxel per axis
        dmax.append(float(dist.max()))
    assert len(set(np.round(dmax, 4))) > 1                                       # the acquisition resolution is re-drawn


def test_training_runs_checkpoints_and_resumes(dataset):
    """tutorial-7 style training() call: 2 epochs x 3 steps, checkpoint per epoch, resume from 002."""
    from SynthSR.training import training
    labels_dir, p, root = dataset
    model_dir = str(root / 'model')
    kw = dict(path_generation_classes=p['classes'], batchsize=1, input_channels=True, output_channel=0, output_shape=32,
              n_levels=3, unet_feat_count=8, steps_per_epoch=3, regression_metric='l1', build_reliability_maps=True,
              work_with_residual_channel=[0], data_res=np.array([1., 1., 2.]), lr=1e-3)
    training(labels_dir, model_dir, p['means'], p['stds'], p['labels'], epochs=2, **kw)
    assert os.path.isfile(os.path.join(model_dir, '001.h5')) and os.path.isfile(os.path.join(model_dir, '002.h5'))
    log = open(os.path.join(model_dir, 'logs', 'loss.csv')).read().strip().splitlines()
    losses = [float(l.split(',')[1]) for l in log]
    assert len(losses) == 2 and all(np.isfinite(losses)) and all(0 < l < 10 for l in losses)
    # the checkpoint is a Keras ModelCheckpoint-style HDF5 file: weights under /model_weights by Keras layer name
    from synthsr_b200 import h5lite
    ck, attrs = h5lite.load_keras_weights(os.path.join(model_dir, '002.h5'))
    assert ck['unet_conv_downarm_0_0/kernel'].shape == (3, 3, 3, 2, 8) and 'unet_likelihood/kernel' in ck
    assert 'unet_maxpool_0' in [n.decode() for n in attrs['layer_names']]
    assert int(h5lite.load_extra(os.path.join(model_dir, '002.h5'))['iterations'][0]) == 6
    training(labels_dir, model_dir, p['means'], p['stds'], p['labels'], epochs=3,
             checkpoint=os.path.join(model_dir, '002.h5'), **kw)
    assert os.path.isfile(os.path.join(model_dir, '003.h5'))
    assert int(h5lite.load_extra(os.path.join(model_dir, '003.h5'))['iterations'][0]) == 9


def test_two_channel_synthesis_config(dataset):
    """multi-modal inputs (T1+T2-like) with a separate HR target channel, batch 2 (c4-like wiring)."""
    from SynthSR.training import training
    labels_dir, p, root = dataset
    pm3 = np.concatenate([np.load(p['means']), np.load(p['means2'])])
    ps3 = np.concatenate([np.load(p['stds']), np.load(p['stds2'])])
    np.save(str(root / 'm3.npy'), pm3); np.save(str(root / 's3.npy'), ps3)
    model_dir = str(root / 'model2')
    training(labels_dir, model_dir, str(root / 'm3.npy'), str(root / 's3.npy'), p['labels'],
             path_generation_classes=p['classes'], batchsize=2, input_channels=[False, True, True], output_channel=0,
             output_shape=32, data_res=np.array([[1., 1., 3.], [1., 1., 2.]]), thickness=np.array([[1., 1., 3.], [1., 1., 2.]]),
             downsample=True, build_reliability_maps=False, n_levels=3, unet_feat_count=8, epochs=1, steps_per_epoch=2)
    from synthsr_b200 import h5lite
    ck, _ = h5lite.load_keras_weights(os.path.join(model_dir, '001.h5'))
    assert ck['unet_conv_downarm_0_0/kernel'].shape == (3, 3, 3, 2, 8)


def test_pipelined_steps_train_the_same_batches_in_order():
    """TrainingEngine.train_step_pipelined (generator of batch i overlaps the U-Net step of batch i-1, second generator
    instance + generator stream) == train_step called on the same batches: same losses (shifted by one call), same
    parameters after flush()."""
    import torch
    from synthsr_b200.generator import GeneratorPlan
    from synthsr_b200.synthetic import GEN_CLASSES, GEN_LABELS, phantom_labels, synthetic_priors
    from synthsr_b200.trainer import TrainingEngine
    shape = [32, 32, 32]
    plan = GeneratorPlan(shape, True, 0, GEN_LABELS, None, 1., None, output_div_by_n=8, translation_bounds=3)
    pm, ps = synthetic_priors(int(GEN_CLASSES.max()) + 1, 1, 0)
    rng = np.random.default_rng(3)
    batches = []
    for i in range(5):
        lab = torch.from_numpy(phantom_labels(shape, GEN_LABELS, seed=i)[None].astype(np.int32))
        m = np.clip(rng.normal(pm[0], pm[1]), 0, None)[GEN_CLASSES][None, :, None].astype(np.float32)
        s = np.clip(rng.normal(ps[0], ps[1]), 0, None)[GEN_CLASSES][None, :, None].astype(np.float32)
        batches.append((lab, m, s))
    res = []
    for pipelined in (False, True):
        eng = TrainingEngine(plan, batchsize=1, nb_levels=3, nb_features=8, conv_impl='ref', seed=5, lr=1e-3)
        losses = []
        for i, (lab, m, s) in enumerate(batches):
            if pipelined:                      # odd batches arrive as pinned host tensors (copied on the generator stream)
                l = eng.train_step_pipelined(lab.pin_memory() if i % 2 else lab.cuda(), m, s)
            else:
                l = eng.train_step(lab.cuda(), m, s)
            if l is not None:
                losses.append(l.item())
        if pipelined:
            losses.append(eng.flush().item())
            assert eng.flush() is None
        torch.cuda.synchronize()
        assert eng.steps == len(batches)
        res.append((losses, eng.net.params.clone()))
    (l_seq, p_seq), (l_pipe, p_pipe) = res
    assert len(l_seq) == len(l_pipe) == 5
    # same kernels on the same batches; what differs run to run is the order of the float atomics in the weight / bias
    # gradients, which Adam's g / sqrt(v) normalisation amplifies for the smallest gradients
    np.testing.assert_allclose(l_pipe, l_seq, rtol=1e-4)
    assert (p_seq - p_pipe).abs().max().item() <= 1e-4 * p_seq.abs().max().item()
