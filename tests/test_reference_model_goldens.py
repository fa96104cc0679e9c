"""The oracle's labels_to_image against the reference's OWN graph-building function executed end to end:
SynthSR/labels_to_image_model.labels_to_image_model() run unmodified on the NumPy `tf` shim with every tf.random draw
logged (tests/golden/make_reference_model_goldens.py -> tests/golden/reference_model.npz).  The oracle is handed the same
inputs and the same draws through the `draws` interface it shares with the product.

Tolerance: the volumes are normalised to [0, 1]; the only arithmetic whose order TF does not specify is the summation
inside tf.nn.conv3d (the shim adds taps in float32, the oracle in float64), so intensities are compared to 2e-6 absolute
(a few float32 ulp); reliability maps and shapes are exact."""
import json
import os

import numpy as np
import pytest

from oracle import generator as OG

HERE = os.path.dirname(os.path.abspath(__file__))
M = np.load(os.path.join(HERE, 'golden', 'reference_model.npz'))
META = json.loads(bytes(M['meta_json']).decode())
ATOL = 2e-6
f32 = np.float32


def _case(tag):
    m = META[tag]
    cfg = dict(m['cfg'])
    cfg['generation_labels'] = M['generation_labels']
    cfg['n_neutral_labels'] = int(M['n_neutral_labels'])
    for k in ('data_res', 'thickness', 'aff'):
        if cfg.get(k) is not None:
            cfg[k] = np.array(cfg[k])
    inputs = [M['%s_in%d' % (tag, i)] for i in range(4) if '%s_in%d' % (tag, i) in M.files]
    pre = '%s_draw_' % tag
    draws = {k[len(pre):]: M[k] for k in M.files if k.startswith(pre)}
    for k in ('aff_rotation', 'aff_shearing', 'aff_scaling', 'aff_translation'):
        draws.setdefault(k, None)
    for k in list(draws):
        if k.startswith('bias_apply'):
            draws[k] = bool(draws[k])
    return cfg, inputs, draws, M['%s_image' % tag], M['%s_target' % tag]


@pytest.mark.parametrize('tag', sorted(META))
def test_whole_graph_matches_reference(tag):
    cfg, inputs, draws, ref_image, ref_target = _case(tag)
    image, target = OG.labels_to_image(cfg, inputs, draws)
    assert image.shape == ref_image.shape and target.shape == ref_target.shape
    assert image.dtype == np.float32 and target.dtype == np.float32
    np.testing.assert_allclose(image, ref_image, rtol=0, atol=ATOL)
    np.testing.assert_allclose(target, ref_target, rtol=0, atol=ATOL)
    assert float(ref_image.std()) > .05 and float(ref_target.std()) > .05          # not a degenerate volume
    if cfg.get('build_reliability_maps'):                                          # maps are exact (0/1 or distances)
        n_in = int(np.sum(cfg['input_channels']))
        rr = cfg.get('randomise_res', False)
        rr = [rr] * len(cfg['input_channels']) if isinstance(rr, bool) else rr
        sim = cfg.get('simulate_registration_error', True)
        first = int(np.argmax(cfg['input_channels']))
        k = 0
        for i, inp in enumerate(cfg['input_channels']):
            if not inp:
                continue
            warped = bool(sim) and i != first                                      # warped maps are interpolated
            if not warped:
                np.testing.assert_array_equal(image[..., 2 * k + 1], ref_image[..., 2 * k + 1])
            k += 1
        assert image.shape[-1] == 2 * n_in


def test_draw_order_of_the_graph_is_the_draws_interface():
    """the number and shapes of the random ops the reference graph executed are exactly what synthsr_b200.draws samples
    (one entry per tf.random call, in graph order) -- case A is training()'s default structure."""
    calls = META['A']['draw_calls']
    kinds = [k for k, _ in calls]
    # rotation, shearing, scaling | svf std, svf | crop | flip | gmm | bias std, bias, bias prob | gamma | blur jitter
    assert kinds == ['uniform'] * 4 + ['normal'] + ['uniform'] * 2 + ['normal'] + ['uniform', 'normal', 'uniform'] + \
        ['normal', 'uniform']
    assert calls[0][1] == [1, 3] and calls[1][1] == [1, 6] and calls[2][1] == [1, 3]
    assert calls[3][1] == [1, 1] and calls[5][1] == [3] and calls[6][1] == [1, 1] and calls[12][1] == [3]


def test_target_rebinding_quirk_is_what_the_reference_does():
    """labels_to_image_model.py:189-196 rebinds `channel` to the blurred, resampled target, so a channel that is both input
    and target at target_res != atlas_res continues on the output grid.  Cases B and E pin it; evaluating the input chain
    from the full-resolution channel instead (the reading one would expect) is measurably different."""
    cfg, inputs, draws, ref_image, _ = _case('B')
    image, _ = OG.labels_to_image(cfg, inputs, draws)
    assert np.abs(image[..., 0] - ref_image[..., 0]).max() < ATOL
    # the "expected" reading: blur(.42 * data_res) of the full-resolution channel, then linear resample
    _, _, inter = OG.labels_to_image(cfg, inputs, draws, return_intermediates=True)
    full = inter['blur_0']
    alt = OG.gaussian_blur(full, [.42, .42, .42], draws['blur_mult_0'])
    alt = OG.resample_tensor(alt[..., None], list(ref_image.shape[1:4]))[..., 0]
    assert np.abs(alt - ref_image[0, ..., 0]).max() > .1


class _EdgeRng:
    """stands in for numpy's Generator inside synthsr_b200.draws.sample_draws: every uniform returns its lower (or upper)
    bound, so the bounds the product samples from can be read off the draws dict."""

    def __init__(self, edge):
        self.edge = edge

    def uniform(self, low=0., high=1., size=None):
        v = np.asarray(low if self.edge == 'lo' else high, dtype=np.float64)
        if size is None:
            size = np.broadcast(np.asarray(low), np.asarray(high)).shape
            if size == ():
                return float(v)
        return np.broadcast_to(v, size).copy()

    def standard_normal(self, size=None, dtype=np.float64):
        return np.ones(size, dtype=dtype)


@pytest.mark.parametrize('tag', sorted(META))
def test_product_samples_from_the_distributions_the_graph_uses(tag):
    """the (minval, maxval) of every tf.random.uniform the reference graph executed (logged by the golden script) against
    the bounds synthsr_b200.draws.sample_draws hands to its generator, entry by entry; normals are standard normals on
    both sides (their scale is applied downstream and is covered by the value comparison above)."""
    from synthsr_b200.draws import sample_draws
    from synthsr_b200.generator import GeneratorPlan
    cfg, inputs, _, _, _ = _case(tag)
    m = META[tag]
    skip = ('input_channels', 'output_channel', 'n_neutral_labels', 'atlas_res', 'target_res', 'generation_labels')
    plan = GeneratorPlan(m['labels_shape'], cfg['input_channels'], cfg['output_channel'], cfg['generation_labels'],
                         cfg['n_neutral_labels'], cfg['atlas_res'], cfg['target_res'],
                         **{k: v for k, v in cfg.items() if k not in skip})
    B = m['batch']
    lo = sample_draws(_EdgeRng('lo'), plan, B, gmm_noise=True)
    hi = sample_draws(_EdgeRng('hi'), plan, B, gmm_noise=True)
    checked = 0
    for key, (kind, a, b) in m['bounds'].items():
        assert key in lo and lo[key] is not None, key
        if kind == 'normal':
            assert np.all(np.asarray(lo[key]) == 1)                                # raw standard normals
            continue
        if key in ('flip',) or key.startswith('bias_apply'):                      # Bernoulli through a U(0,1) threshold
            assert (a, b) == (0., 1.)
            assert np.all(np.asarray(lo[key])) and not np.any(np.asarray(hi[key]))
        elif key == 'crop_idx':
            assert np.all(np.asarray(lo[key]) == 0)
            np.testing.assert_array_equal(np.asarray(hi[key])[0], np.asarray(b).astype(np.int32))
        elif key.startswith('thick_'):                                             # U(atlas_res, resolution drawn just before)
            np.testing.assert_allclose(np.asarray(lo[key]), np.broadcast_to(np.asarray(a, f32), np.shape(lo[key])))
            np.testing.assert_array_equal(np.asarray(hi[key]), np.asarray(hi['res_' + key[6:]]))
        else:
            np.testing.assert_allclose(np.asarray(lo[key], f32), np.broadcast_to(np.asarray(a, f32), np.shape(lo[key])), rtol=1e-6)
            np.testing.assert_allclose(np.asarray(hi[key], f32), np.broadcast_to(np.asarray(b, f32), np.shape(hi[key])), rtol=1e-6)
        checked += 1
    assert checked >= 5
    # and nothing is sampled that the graph does not draw
    drawn = {k for k, v in lo.items() if v is not None and not (k == 'crop_idx' and 'crop_idx' not in m['bounds'])
             and not (k == 'flip' and 'flip' not in m['bounds'])}
    assert drawn == set(m['bounds']), drawn ^ set(m['bounds'])
