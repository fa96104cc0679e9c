"""Drop-in check of the API surface: the reference's OWN user scripts -- scripts/tutorials/1..7 and the scripts/training.py
command line -- are executed UNMODIFIED (runpy, copied to a scratch tree next to a copy of the reference's data/ folder, never
into this repository) against THIS package's `SynthSR` / `ext` modules.  The CUDA engines are replaced by recorders (this is
the CPU suite), so what is checked is everything a user's script touches: constructor keywords and defaults, attributes
(`brain_generator.aff`, `.header`), the shapes and channel counts `generate_brain()` returns for each tutorial configuration,
the NIfTI files the tutorials write, and what `training()` derives from tutorial 7 and from the command line.

Runs only where /root/reference is mounted (the build container); skipped on the GPU box."""
import os
import runpy
import shutil
import sys

import numpy as np
import pytest

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'scripts', 'tutorials')),
                                reason='reference tree not mounted')
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _StubGenerator:
    """stands in for synthsr_b200.generator.SynthGenerator: checks what the host side hands over, returns volumes of the
    planned shapes (a smooth function of the deformed-label input so that files are not constant)."""
    made = []

    def __init__(self, plan, batchsize=1, device='cuda'):
        import torch
        self.plan, self.B, self.device = plan, int(batchsize), torch.device('cpu')
        _StubGenerator.made.append(self)
        self.runs = 0

    def run(self, labels, means, stds, draws, real_image=None, seed=0, keep=None):
        import torch
        p, B = self.plan, self.B
        assert labels.dtype == torch.int32 and list(labels.shape) == [B] + p.labels_shape
        assert np.asarray(means).shape == (B, len(p.generation_labels), p.n_channels) == np.asarray(stds).shape
        assert (real_image is not None) == p.use_real_image
        for k in ('crop_idx', 'flip'):
            assert k in draws
        self.runs += 1
        g = np.linspace(0, 1, int(np.prod(p.output_shape))).reshape(p.output_shape).astype(np.float32)
        image = np.stack([g * (c + 1) for c in range(p.n_image_channels)], -1)[None].repeat(B, 0)
        target = np.stack([1 - g for _ in range(p.n_target_channels)], -1)[None].repeat(B, 0)
        return torch.from_numpy(image), torch.from_numpy(target)


class _RecordedEngine:
    last = None

    def __init__(self, plan, **kw):
        self.plan, self.kw, self.net = plan, kw, None
        _RecordedEngine.last = self


@pytest.fixture
def scratch(tmp_path, monkeypatch):
    """<tmp>/scripts/tutorials/*.py, <tmp>/scripts/training.py and <tmp>/data/{labels,images,labels_classes_priors} (links to the
    reference's files; results are written to new folders under <tmp>/data, never through a link)."""
    (tmp_path / 'scripts' / 'tutorials').mkdir(parents=True)
    for f in os.listdir(os.path.join(REF, 'scripts', 'tutorials')):
        shutil.copy(os.path.join(REF, 'scripts', 'tutorials', f), str(tmp_path / 'scripts' / 'tutorials' / f))
    shutil.copy(os.path.join(REF, 'scripts', 'training.py'), str(tmp_path / 'scripts' / 'training.py'))
    for d in ('labels', 'images', 'labels_classes_priors'):
        (tmp_path / 'data' / d).mkdir(parents=True)
        for f in os.listdir(os.path.join(REF, 'data', d)):
            os.symlink(os.path.join(REF, 'data', d, f), str(tmp_path / 'data' / d / f))
    import synthsr_b200.generator as G
    import synthsr_b200.trainer as TR
    import SynthSR.training as PT
    monkeypatch.setattr(G, 'SynthGenerator', _StubGenerator)
    monkeypatch.setattr(TR, 'TrainingEngine', _RecordedEngine)
    monkeypatch.setattr(PT, 'train_model', lambda *a, **k: setattr(_RecordedEngine, 'train_args', (a, k)))
    monkeypatch.setattr(PT, 'metrics_model', lambda *a, **k: None)
    monkeypatch.setattr(PT.nrn_models, 'UnetModel', lambda *a, **k: None)
    monkeypatch.chdir(str(tmp_path / 'scripts' / 'tutorials'))
    monkeypatch.setattr(sys, 'path', [REPO] + [p for p in sys.path if not p.startswith(REF)])
    _StubGenerator.made.clear()
    return tmp_path


#            script                                   image channels  target ch  output shape (None: label-map shape)
TUTORIALS = [('1-SR_real.py', 2, 1, None),
             ('2-SR_synthetic.py', 2, 1, [98, 124, 103]),      # min(int(148, 187, 155 / 1.5), 128)  (get_shapes)
             ('3-synthesis_real.py', 2, 1, [128, 128, 128]),
             ('4-synthesis_synthetic.py', 2, 1, [128, 128, 128]),
             ('5-SR-synthesis_multimodal_real.py', 4, 1, [128, 128, 128]),
             ('6-SR-synthesis_multimodal_synthetic.py', 4, 1, [128, 128, 128])]


def test_generation_tutorials_fail_on_the_same_missing_keyword_as_with_the_reference(scratch):
    """tutorials 1-6 predate BrainGenerator's required `prior_distributions` argument (brain_generator.py:34): against the
    reference's own class they raise TypeError, and so they must here."""
    import inspect
    import json
    sig = json.load(open(os.path.join(REPO, 'tests', 'golden', 'reference_signatures.json')))
    ref_params = sig['SynthSR/brain_generator.py:BrainGenerator.__init__']
    assert any(n == 'prior_distributions' and d is None for n, d in ref_params)          # required in the reference
    from SynthSR.brain_generator import BrainGenerator
    assert inspect.signature(BrainGenerator.__init__).parameters['prior_distributions'].default is inspect.Parameter.empty
    with pytest.raises(TypeError, match="missing 1 required positional argument: 'prior_distributions'"):
        runpy.run_path('1-SR_real.py', run_name='__main__')


@pytest.mark.parametrize('script,n_img,n_tgt,out_shape', TUTORIALS, ids=[t[0][:-3] for t in TUTORIALS])
def test_generation_tutorial_runs_unmodified(scratch, monkeypatch, script, n_img, n_tgt, out_shape):
    """... with the one keyword the tutorials forget defaulted to 'normal' (training()'s default, and what the tutorials'
    (2, K) mean / std prior files are made for)."""
    from ext.lab2im import utils
    import SynthSR.brain_generator as BG

    class WithDefault(BG.BrainGenerator):
        def __init__(self, *a, **k):
            k.setdefault('prior_distributions', 'normal')
            super().__init__(*a, **k)

    monkeypatch.setattr(BG, 'BrainGenerator', WithDefault)
    ns = runpy.run_path(script, run_name='__main__')
    gen = _StubGenerator.made[-1]
    p = gen.plan
    assert gen.runs == ns['n_examples'] == 3
    assert p.n_image_channels == n_img and p.n_target_channels == n_tgt
    labels_shape = [148, 187, 155]
    if out_shape is None:                                     # tutorial 1: no cropping -> largest shape divisible by nothing
        out_shape = labels_shape
    assert p.output_shape == out_shape, (p.output_shape, out_shape)
    assert p.labels_shape == labels_shape and p.use_real_image == (ns['output_channel'] is None)
    bg = ns['brain_generator']
    assert bg.aff.shape == (4, 4) and list(bg.labels_shape) == labels_shape
    files = sorted(os.listdir(ns['result_dir']))
    assert len(files) == 3 * (n_img + n_tgt), files
    vol = utils.load_volume(os.path.join(ns['result_dir'], files[0]))
    assert sorted(vol.shape) == sorted(out_shape)
    if script.startswith('2-'):
        # tutorial 2 is the configuration where the target channel is also the input at target_res 1.5: its acquisition chain
        # runs on the output grid (GeneratorPlan.chan_grid; labels_to_image_model.py:193-195)
        assert p.crop_shape != p.output_shape and p.chan_grid == [p.output_shape]
        assert p.crop_shape == [147, 186, 154] and p.down_shape == [[98, 124, int(103 / 3)]]


def test_training_tutorial_runs_unmodified(scratch):
    ns = runpy.run_path('7-training.py', run_name='__main__')
    eng = _RecordedEngine.last
    p = eng.plan
    assert p.input_channels == [False, True, True] and p.output_channel == [0] and p.output_shape == [128, 128, 128]
    assert p.n_image_channels == 4 and p.build_reliability_maps and p.sim_reg == [True, True, True]
    assert eng.kw['nb_levels'] == ns['n_levels'] == 5 and eng.kw['nb_features'] == ns['unet_feat_count']
    assert eng.kw['metric'] == ns['regression_metric'] and eng.kw['lr'] == ns['learning_rate']
    # work_with_residual_channel = 1 with reliability maps: the reference repeats the list ([1, 1]) -> image_out channel 1
    assert eng.kw['work_with_residual_channel'] == [1]
    assert eng.kw['loss_cropping'] is None
    a, k = _RecordedEngine.train_args
    assert a[4] == ns['epochs'] and a[5] == ns['steps_per_epoch'] and os.path.isdir(ns['model_dir'])


def test_training_command_line_runs_unmodified(scratch, monkeypatch):
    """scripts/training.py of the reference: argparse -> training(**vars(args)) (uses ext.lab2im.utils.infer for bounds)."""
    d = str(scratch / 'data')
    np.save(str(scratch / 'res.npy'), np.array([1., 1., 4.]))
    argv = ['training.py', d + '/labels', str(scratch / 'cli_models'), d + '/labels_classes_priors/prior_means_t1_lr.npy',
            d + '/labels_classes_priors/prior_stds_t1_lr.npy', d + '/labels_classes_priors/generation_labels.npy',
            '--generation_classes', d + '/labels_classes_priors/generation_classes.npy', '--output_channel', '0',
            '--output_shape', '96', '--rotation', '10', '--translation', 'False', '--data_res', str(scratch / 'res.npy'), '--no_rel_map',
            '--n_levels', '4', '--unet_feat', '16', '--lr', '2e-4', '--epochs', '3', '--steps_per_epoch', '7',
            '--metric', 'l2', '--loss_cropping', '64']
    monkeypatch.setattr(sys, 'argv', argv)
    runpy.run_path(str(scratch / 'scripts' / 'training.py'), run_name='__main__')
    eng = _RecordedEngine.last
    p = eng.plan
    assert p.output_shape == [96, 96, 96] and p.rotation_bounds == 10 and p.translation_bounds is False
    assert not p.build_reliability_maps and p.n_image_channels == 1
    np.testing.assert_array_equal(p.data_res, [[1., 1., 4.]])
    assert eng.kw['nb_levels'] == 4 and eng.kw['nb_features'] == 16 and eng.kw['lr'] == 2e-4 and eng.kw['metric'] == 'l2'
    assert eng.kw['loss_cropping'] == 64
    a, k = _RecordedEngine.train_args
    assert a[4] == 3 and a[5] == 7
