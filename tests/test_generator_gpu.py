"""GPU parity: CUDA generator (through the C ABI) vs the NumPy oracle on identical inputs and injected draws.

Bars: label resampling bit exact; integrated deformation field bit exact (same float32 op order, no FMA);
float images within 1e-3 of the image range (values are min-max normalised to [0,1]); in practice ~1e-6.
"""
import numpy as np
import pytest
import torch

from helpers import GEN_LABELS, SIDED_LABELS, gmm_params, phantom_labels

pytestmark = pytest.mark.gpu

TOL = 1e-3


def _run_case(cfg, labels_shape, label_list, seed=0, batch=1, real=False, phantom=phantom_labels):
    from oracle import generator as OG
    from synthsr_b200.draws import sample_draws
    from synthsr_b200.generator import GeneratorPlan, SynthGenerator

    rng = np.random.default_rng(seed)
    labs = np.stack([phantom(labels_shape, label_list, seed=seed + b) for b in range(batch)])
    plan = GeneratorPlan(labels_shape, cfg.get('input_channels', True), cfg.get('output_channel', 0), label_list,
                         cfg.get('n_neutral_labels'), cfg.get('atlas_res', 1.), cfg.get('target_res'),
                         **{k: v for k, v in cfg.items() if k not in ('input_channels', 'output_channel',
                                                                      'n_neutral_labels', 'atlas_res', 'target_res')})
    means, stds = gmm_params(rng, len(label_list), plan.n_channels, batch)
    draws = sample_draws(rng, plan, batch, gmm_noise=True)
    gen = SynthGenerator(plan, batchsize=batch)
    keep = {}
    inputs = [labs[..., None], means, stds]
    real_t = None
    if real:
        real_np = rng.uniform(0, 200, size=(batch, *labels_shape, 1)).astype(np.float32)
        inputs.append(real_np)
        real_t = torch.from_numpy(real_np[..., 0]).cuda().contiguous()
    image, target = gen.run(torch.from_numpy(labs).cuda(), means, stds, draws, real_image=real_t, keep=keep)
    torch.cuda.synchronize()
    ocfg = dict(cfg)
    ocfg['generation_labels'] = label_list
    o_image, o_target, inter = OG.labels_to_image(ocfg, inputs, draws, return_intermediates=True)
    return plan, image.cpu().numpy(), target.cpu().numpy(), keep, o_image, o_target, inter


def _check(plan, image, target, keep, o_image, o_target, inter):
    assert image.shape == o_image.shape and target.shape == o_target.shape
    if 'integrated' in keep:
        np.testing.assert_array_equal(keep['integrated'][0].cpu().numpy(), inter['integrated'])
    np.testing.assert_array_equal(keep['labels'][0].cpu().numpy(), inter['labels'])          # bit exact
    assert np.abs(image - o_image).max() <= TOL, np.abs(image - o_image).max()
    assert np.abs(target - o_target).max() <= TOL, np.abs(target - o_target).max()


def test_default_single_channel_crop():
    """BrainGenerator defaults (brain_generator.py:30-61) on a 40x48x36 phantom cropped to 32^3."""
    cfg = dict(output_shape=32, translation_bounds=5, aff=np.eye(4))
    _check(*_run_case(cfg, [40, 48, 36], GEN_LABELS, seed=1))


def test_training_defaults_no_crop_batch2():
    """training() defaults (training.py:57-73): factor .03125, nonlin 4, shear .02, no cropping, batch 2."""
    cfg = dict(scaling_bounds=.15, rotation_bounds=15, shearing_bounds=.02, translation_bounds=5, nonlin_std=4.,
               nonlin_shape_factor=.03125, bias_field_std=.3, bias_shape_factor=.03125, output_div_by_n=32)
    _check(*_run_case(cfg, [64, 64, 32], GEN_LABELS, seed=2, batch=2))


def test_benchmark_shape_160_training_defaults():
    """BASELINE configs[1]/[2]: the 160^3 label map of bench.py through the generator with training()'s defaults
    (SynthSR/training.py:57-73), against the oracle at full size: integrated SVF and labels bit exact, images <= 1e-3."""
    import bench
    cfg = {k: v for k, v in bench.TRAINING_DEFAULTS.items()}
    for seed in (0, 1):
        _check(*_run_case(cfg, [160, 160, 160], GEN_LABELS, seed=seed, phantom=lambda shape, labels, seed: bench.make_inputs(
            shape, 1, seed=seed)[0][0]))


def test_benchmark_shape_64_brain_generator_defaults():
    """BASELINE configs[0]: single 64^3 label map with BrainGenerator's own defaults (SynthSR/brain_generator.py:30-61:
    nonlin_std 3, factor .0625, shearing .012, bias factor .025, no reliability maps), every intermediate the oracle keeps."""
    for seed in (0, 1, 2):
        plan, image, target, keep, o_image, o_target, inter = _run_case(dict(aff=np.eye(4)), [64, 64, 64], GEN_LABELS, seed=seed)
        _check(plan, image, target, keep, o_image, o_target, inter)
        assert image.shape == (1, 64, 64, 64, 1) and plan.svf_small_shape == [4, 4, 4] and plan.svf_half_shape == [32, 32, 32]
        for name in ('raw_0', 'blur_0'):
            if name in inter:
                scale = max(1., float(np.abs(inter[name]).max()))
                assert np.abs(keep[name].cpu().numpy().reshape(inter[name].shape) - inter[name]).max() <= TOL * scale, name


def test_flip_swaps_sided_labels():
    cfg = dict(output_shape=32, n_neutral_labels=4)
    from synthsr_b200.generator import GeneratorPlan  # noqa: F401
    for seed in (3, 4, 5, 6):   # both flip outcomes are exercised over the seeds
        _check(*_run_case(cfg, [36, 40, 34], SIDED_LABELS, seed=seed))


def test_multichannel_lowres_regerror_relmaps():
    """Hyperfine-like: HR target + two LR inputs, thick slices, downsampling, registration error, reliability maps
    (labels_to_image_model.py:85-92, 202-238)."""
    cfg = dict(input_channels=[False, True, True], output_channel=0, data_res=np.array([[1., 1., 3.], [1., 1., 4.]]),
               thickness=np.array([[1., 1., 2.], [1., 1., 4.]]), downsample=True, build_reliability_maps=True,
               simulate_registration_error=True, output_shape=32)
    _check(*_run_case(cfg, [40, 40, 40], GEN_LABELS, seed=7))


def test_randomise_res_single_channel():
    """randomise_res branch (labels_to_image_model.py:215-220): SampleResolution draws -> per-example separable
    DynamicGaussianBlur -> fused MimicAcquisition, with the acquisition distance map as reliability channel."""
    cfg = dict(randomise_res=True, build_reliability_maps=True, output_shape=32)
    plan, image, target, keep, o_image, o_target, inter = _run_case(cfg, [40, 44, 36], GEN_LABELS, seed=11, batch=2)
    assert plan.n_image_channels == 2
    _check(plan, image, target, keep, o_image, o_target, inter)
    assert np.abs(image[..., 1] - o_image[..., 1]).max() <= 1e-5          # distance map: same float32 operations


def test_randomise_res_two_channels_regerror():
    """randomised acquisition on two synthetic inputs with registration error on the second (+ warped distance maps)."""
    cfg = dict(input_channels=[True, True], output_channel=0, randomise_res=True, build_reliability_maps=True,
               simulate_registration_error=True, output_shape=32)
    _check(*_run_case(cfg, [40, 40, 40], GEN_LABELS, seed=12))


def test_target_resampling_and_padding():
    """target_res != atlas_res (blur + linear resample of the target, :189-196) and padding_margin (:116-120)."""
    cfg = dict(target_res=2., padding_margin=4, nonlin_std=2.)
    _check(*_run_case(cfg, [32, 40, 32], GEN_LABELS, seed=8))


def test_no_deformation_identity_labels():
    """all spatial augmentation off => labels pass through unchanged (analytic KAT)."""
    cfg = dict(scaling_bounds=False, rotation_bounds=False, shearing_bounds=False, translation_bounds=False,
               nonlin_std=0., flipping=False)
    plan, image, target, keep, o_image, o_target, inter = _run_case(cfg, [32, 32, 32], GEN_LABELS, seed=9)
    labs = phantom_labels([32, 32, 32], GEN_LABELS, seed=9)
    np.testing.assert_array_equal(keep['labels'][0].cpu().numpy(), labs)
    _check(plan, image, target, keep, o_image, o_target, inter)


def test_real_image_target():
    """images_dir path: real image deformed with linear interpolation, normalised target (:109-134, 248-255)."""
    cfg = dict(output_channel=None, output_shape=32)
    _check(*_run_case(cfg, [40, 40, 36], GEN_LABELS, seed=10, real=True))


def test_philox_noise_moments():
    """throughput mode: on-device Philox normals; a single-label map must give image ~ N(mu, sigma) before
    normalisation (statistical KAT)."""
    from synthsr_b200._lib import lib, stream_ptr
    n = 1 << 20
    out = torch.empty(n, dtype=torch.float32, device='cuda')
    lib.ssr_philox_normal(out, n, 1234, 7, stream_ptr())
    torch.cuda.synchronize()
    x = out.double()
    assert abs(x.mean().item()) < 5e-3 and abs(x.std().item() - 1) < 5e-3
    assert abs(((x ** 3).mean()).item()) < 2e-2 and abs(((x ** 4).mean()).item() - 3) < 5e-2
    out2 = torch.empty(n, dtype=torch.float32, device='cuda')
    lib.ssr_philox_normal(out2, n, 1234, 8, stream_ptr())
    assert abs(torch.corrcoef(torch.stack([out, out2]))[0, 1].item()) < 5e-3
