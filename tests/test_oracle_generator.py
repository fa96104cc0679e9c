"""CPU tests pinning the NumPy oracle against the reference's docstring worked examples and analytic known-answer
cases (the reference ships no tests or golden vectors -- SURVEY.md 4 / 8c), and the product's host-side plan / draws
logic against the oracle's independent restatement."""
import numpy as np

from helpers import GEN_LABELS, phantom_labels
from oracle import generator as OG

f32 = np.float32


def _identity_draws(shape, batch=1, C=1):
    return {'aff_rotation': np.zeros((batch, 3), f32), 'aff_shearing': np.zeros((batch, 6), f32),
            'aff_scaling': np.ones((batch, 3), f32), 'aff_translation': np.zeros((batch, 3), f32),
            'crop_idx': np.zeros((batch, 3), np.int32), 'flip': np.zeros(batch, bool),
            'gmm_normal': np.zeros((batch, *shape, C), f32), 'gamma_normal_0': np.zeros(batch, f32),
            'bias_apply_0': False, 'bias_std_0': np.zeros(batch, f32), 'bias_normal_0': np.zeros((batch, 1, 1, 1), f32),
            'blur_mult_0': np.ones(3, f32)}


def test_identity_affine_zero_field_keeps_labels():
    shape = [20, 24, 18]
    lab = phantom_labels(shape, GEN_LABELS, seed=0)
    shift = OG.affine_to_shift(np.eye(4, dtype=f32), shape, np.zeros(shape + [3], f32))
    assert np.all(shift == 0)
    out = OG.transform(lab.astype(f32)[..., None], shift, 'nearest')[..., 0]
    np.testing.assert_array_equal(out.astype(np.int32), lab)


def test_integer_translation_shifts_with_edge_clamp():
    shape = [10, 11, 12]
    vol = np.arange(np.prod(shape), dtype=f32).reshape(shape)[..., None]
    aff = np.eye(4, dtype=f32)
    aff[:3, 3] = [2, -1, 3]
    out = OG.spatial_transformer(vol, aff, None, 'nearest')[..., 0]
    i, j, k = np.meshgrid(*[np.arange(s) for s in shape], indexing='ij')
    exp = vol[np.clip(i + 2, 0, 9), np.clip(j - 1, 0, 10), np.clip(k + 3, 0, 11), 0]
    np.testing.assert_array_equal(out, exp)


def test_round_half_to_even():
    vol = np.arange(6, dtype=f32).reshape(6, 1, 1, 1)
    loc = [np.array([0.5, 1.5, 2.5, 3.5], f32), np.zeros(4, f32), np.zeros(4, f32)]
    np.testing.assert_array_equal(OG.interpn_nearest(vol, loc)[:, 0], [0, 2, 2, 4])     # tf.round semantics


def test_linear_interp_edge_clamp_and_midpoint():
    vol = np.array([1., 3., 7.], f32).reshape(3, 1, 1, 1)
    loc = [np.array([-2., 0.5, 1.25, 2., 5.], f32), np.zeros(5, f32), np.zeros(5, f32)]
    np.testing.assert_allclose(OG.interpn_linear(vol, loc)[:, 0], [1., 2., 4., 7., 7.], rtol=0, atol=1e-6)


def test_resize_same_shape_is_identity():
    rng = np.random.default_rng(0)
    v = rng.normal(size=(5, 6, 7, 3)).astype(f32)
    np.testing.assert_array_equal(OG.resize(v, [5, 6, 7]), v)


def test_resize_samples_at_j_in_over_out():
    v = np.arange(4, dtype=f32).reshape(4, 1, 1, 1)
    out = OG.resize(v, [8, 1, 1])[:, 0, 0, 0]
    np.testing.assert_allclose(out, np.minimum(np.arange(8) * 0.5, 3.0), atol=1e-6)    # origin aligned, clamped


def test_vecint_constant_field():
    """VecInt of a constant field c is c (every step composes c/2^k with itself; edge clamped)."""
    v = np.zeros((8, 8, 8, 3), f32)
    v[..., 0] = 1.5
    v[..., 2] = -0.75
    out = OG.integrate_vec(v, 7)
    np.testing.assert_allclose(out, v, atol=1e-5)


def test_random_flip_docstring_example():
    """ext/lab2im/layers.py:306-320 (RandomFlip example 3): flipping + swapping labels 1 <-> 2."""
    inp = np.array([[1, 0, 0, 0, 0, 0, 0], [1, 0, 0, 0, 2, 2, 0], [1, 0, 0, 0, 2, 2, 0], [1, 0, 0, 0, 2, 2, 0],
                    [1, 0, 0, 0, 0, 0, 0]])
    exp = np.array([[0, 0, 0, 0, 0, 0, 2], [0, 1, 1, 0, 0, 0, 2], [0, 1, 1, 0, 0, 0, 2], [0, 1, 1, 0, 0, 0, 2],
                    [0, 0, 0, 0, 0, 0, 2]])
    label_list, n_neutral = np.array([0, 1, 2]), 1
    split = np.split(label_list, [n_neutral, n_neutral + 1])
    lut = OG.get_mapping_lut(label_list, np.concatenate((split[0], split[2], split[1])))
    np.testing.assert_array_equal(lut[inp][:, ::-1], exp)
    # same through the full graph: volume [7 (flip axis 0), 5, 1]
    vol = np.transpose(inp)[..., None]
    cfg = dict(generation_labels=label_list, n_neutral_labels=1, scaling_bounds=False, rotation_bounds=False,
               shearing_bounds=False, translation_bounds=False, nonlin_std=0., bias_field_std=0.)
    d = _identity_draws([7, 5, 1])
    d['flip'] = np.array([True])
    _, _, inter = OG.labels_to_image(cfg, [vol[None, ..., None], np.zeros((1, 3, 1), f32), np.zeros((1, 3, 1), f32)], d,
                                     return_intermediates=True)
    np.testing.assert_array_equal(inter['labels'][..., 0], np.transpose(exp))


def test_get_dims_and_shapes_table():
    """SURVEY appendix A shape table (probed with the reference formulas)."""
    assert OG.get_resample_shape([148, 187, 155], .03125) == [5, 6, 5]
    assert OG.get_resample_shape([148, 187, 155], .0625) == [10, 12, 10]
    assert OG.get_resample_shape([160] * 3, .03125) == [5, 5, 5]
    assert OG.get_resample_shape([160] * 3, .025) == [4, 4, 4]
    assert OG.get_resample_shape([256] * 3, .025) == [7, 7, 7]
    assert OG.get_resample_shape([192, 192, 64], .025) == [5, 5, 2]
    crop, out, _ = OG.get_shapes([148, 187, 155], None, [1.] * 3, [1.] * 3, None, 32)
    assert crop == [128, 160, 128] and out == crop
    crop, out, _ = OG.get_shapes([148, 187, 155], 128, [1.] * 3, [1.] * 3, None, 32)
    assert crop == [128] * 3 and out == crop
    assert OG.find_closest_number_divisible_by_m(187, 32) == 160        # utils.py:928-944 'lower'


def test_gaussian_kernel_values():
    """sigma=.5 -> 3^3 window, 1-D normalised taps [0.106507, 0.786986, 0.106507]."""
    k = OG.gaussian_kernel_dense([.5, .5, .5])
    assert k.shape == (3, 3, 3) and abs(k.sum() - 1) < 1e-6
    np.testing.assert_allclose(k.sum((1, 2)), [0.106507, 0.786986, 0.106507], atol=1e-5)
    assert OG.gaussian_kernel_dense([.42, .42, 1.26]).shape == (3, 3, 5)
    assert OG.gaussian_kernel_dense(list(.42 * np.array([1.5, 1.5, 5.]))).shape == (3, 3, 7)
    ks = OG.gaussian_kernels_separable([6., 0.3, 0.3])
    assert len(ks[0]) == 15 and ks[1] is None


def test_blur_of_constant_darkens_edges_only():
    img = np.ones((6, 6, 6), f32)
    out = OG.gaussian_blur(img, .5)
    assert abs(out[3, 3, 3] - 1) < 1e-6 and out[0, 0, 0] < 0.75 and out[0, 3, 3] < 0.9   # zero padding ('SAME')


def test_reliability_map_example():
    """out=128, down=42 tent weights."""
    rel = OG.reliability_map([128, 4, 4], [42, 4, 4])[:, 0, 0]
    np.testing.assert_allclose(rel[:10], [1, 0, 0, .952381, .047619, 0, .904762, .095238, 0, .857143], atol=1e-5)


def test_single_label_gmm_moments():
    shape = [24, 24, 24]
    lab = np.full(shape, 14, np.int32)
    rng = np.random.default_rng(0)
    cfg = dict(generation_labels=GEN_LABELS, scaling_bounds=False, rotation_bounds=False, shearing_bounds=False,
               translation_bounds=False, nonlin_std=0., bias_field_std=0., flipping=False)
    d = _identity_draws(shape)
    d['gmm_normal'] = rng.standard_normal((1, *shape, 1)).astype(f32)
    means = np.zeros((1, len(GEN_LABELS), 1), f32)
    stds = np.zeros_like(means)
    means[0, 1], stds[0, 1] = 100., 10.
    _, _, inter = OG.labels_to_image(cfg, [lab[None, ..., None], means, stds], d, return_intermediates=True)
    g = inter['gmm'][..., 0]
    assert abs(g.mean() - 100) < 0.3 and abs(g.std() - 10) < 0.3


def test_plan_matches_oracle_bookkeeping():
    """host logic of the product (GeneratorPlan) vs the oracle's independent restatement of
    labels_to_image_model.py:69-100."""
    from synthsr_b200.generator import GeneratorPlan, gaussian_kernel, reliability_factors
    cfgs = [
        (dict(output_shape=32), [40, 48, 36]),
        (dict(input_channels=[False, True, True], output_channel=0, data_res=np.array([[1., 1., 3.], [1., 1., 4.]]),
              thickness=np.array([[1., 1., 2.], [1., 1., 4.]]), downsample=True, output_shape=32), [40, 40, 40]),
        (dict(target_res=2., padding_margin=4), [32, 40, 32]),
        (dict(output_div_by_n=32), [148, 187, 155]),
        (dict(output_channel=None), [40, 40, 36]),
    ]
    for cfg, shape in cfgs:
        r = OG.resolve_config(cfg, shape)
        p = GeneratorPlan(shape, cfg.get('input_channels', True), cfg.get('output_channel', 0), GEN_LABELS, None,
                          cfg.get('atlas_res', 1.), cfg.get('target_res'),
                          **{k: v for k, v in cfg.items() if k not in ('input_channels', 'output_channel', 'target_res')})
        assert p.crop_shape == r['crop_shape'] and p.output_shape == r['output_shape']
        assert p.input_channels == r['input_channels'] and p.output_channel == r['output_channel']
        np.testing.assert_array_equal(p.data_res, r['data_res'])
        np.testing.assert_array_equal(p.thickness, r['thickness'])
        assert [bool(v) for v in p.downsample] == [bool(v) for v in r['downsample']]
    k = gaussian_kernel([.5, .5, .5])[0]
    np.testing.assert_array_equal(k, OG.gaussian_kernel_dense([.5, .5, .5]))
    m = np.array([1.1, .9, 1.05], f32)
    np.testing.assert_array_equal(gaussian_kernel([.42, .42, 1.26], m)[0], OG.gaussian_kernel_dense([.42, .42, 1.26], m))
    f = reliability_factors([128, 8, 8], [42, 8, 8])
    rel = (f[0][:, None, None] * f[1][None, :, None] * f[2][None, None, :]).astype(f32)
    np.testing.assert_array_equal(rel, OG.reliability_map([128, 8, 8], [42, 8, 8]))


def test_affine_builders_bit_identical():
    """the 4x4 matrix must be bit identical between product host code and oracle (label output is bit exact)."""
    from synthsr_b200 import draws as D
    rng = np.random.default_rng(3)
    for _ in range(20):
        rot, sh = rng.uniform(-15, 15, 3).astype(f32), rng.uniform(-.02, .02, 6).astype(f32)
        sc, tr = rng.uniform(.85, 1.15, 3).astype(f32), rng.uniform(-5, 5, 3).astype(f32)
        a, b = D.build_affine(rot, sh, sc, tr), OG.build_affine(rot, sh, sc, tr)
        assert a.dtype == f32 and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_sample_draws_shapes_and_ranges():
    from synthsr_b200.draws import sample_draws
    from synthsr_b200.generator import GeneratorPlan
    p = GeneratorPlan([160, 160, 160], True, 0, GEN_LABELS, None, 1., None, translation_bounds=5, nonlin_std=4.,
                      nonlin_shape_factor=.03125, shearing_bounds=.02, bias_shape_factor=.03125)
    assert p.svf_small_shape == [5, 5, 5] and p.svf_half_shape == [80, 80, 80] and p.bias_small_shape == [5, 5, 5]
    d = sample_draws(np.random.default_rng(0), p, 2)
    assert d['svf_normal'].shape == (2, 5, 5, 5, 3) and d['aff_rotation'].shape == (2, 3)
    assert np.all(np.abs(d['aff_rotation']) <= 15) and np.all(np.abs(d['aff_scaling'] - 1) <= .15)
    assert np.all(np.abs(d['aff_shearing']) <= .02) and 0 <= d['svf_std'] <= 4
    assert 'gmm_normal' not in d and d['bias_normal_0'].shape == (2, 5, 5, 5)
    assert np.all((d['blur_mult_0'] >= 1 / 1.15) & (d['blur_mult_0'] <= 1.15))


def test_mimic_acquisition_identity_and_distance_pattern():
    """ext/lab2im/layers.py:921-987: acquisition at the volume resolution is the identity with zero distance; at
    [1,1,3] mm every third plane along the last axis is 'acquired' (distance 0) and the distance peaks in between."""
    rng = np.random.default_rng(0)
    vol = rng.uniform(size=(12, 10, 18, 1)).astype(np.float32)
    out, dist = OG.mimic_acquisition(vol, [1., 1., 1.], [1., 1., 1.], [1., 1., 1.], [12, 10, 18])
    np.testing.assert_array_equal(out, vol)
    assert dist.max() == 0
    out, dist = OG.mimic_acquisition(vol, [1., 1., 3.], [1., 1., 1.], [1., 1., 1.], [12, 10, 18])
    assert out.shape == vol.shape and dist.shape == vol.shape
    np.testing.assert_array_equal(out[:, :, 0], vol[:, :, 0])                # first plane is sampled exactly
    np.testing.assert_allclose(dist[0, 0, :7, 0], [0., 1., 1., 0., 1., 1., 0.], atol=1e-5)   # |j/3 - round(j/3)| * 3 mm
    ds, dz, uz = OG.mimic_acquisition_zooms([12, 10, 18], [1., 1., 1.], [1., 1., 3.], [12, 10, 18])
    assert list(ds) == [12, 10, 6] and np.allclose(dz, [1, 1, 1 / 3]) and np.allclose(uz, [1, 1, 3])


def test_dynamic_kernels_host_logic_matches_oracle():
    """product-side host helpers of the randomise_res branch against the oracle restatement (same float32 formulas),
    including the reference's normalisation of the per-example kernels by the sum over the whole batch."""
    from synthsr_b200.generator import dynamic_kernels, dynamic_sigma, mimic_zooms
    rng = np.random.default_rng(3)
    res = rng.uniform(1., 9., size=(3, 3)).astype(np.float32)
    thick = (1. + rng.uniform(size=(3, 3)) * (res - 1.)).astype(np.float32)
    mult = rng.uniform(1 / 1.15, 1.15, size=(3, 3)).astype(np.float32)
    sig = dynamic_sigma([1., 1., 1.], res, thick)
    np.testing.assert_array_equal(sig, OG.dynamic_sigma([1., 1., 1.], res, thick))
    np.testing.assert_allclose(sig, 0.42 * np.minimum(res, thick), rtol=1e-6)
    ks = dynamic_kernels(sig, [17, 17, 17], mult)
    oks = OG.dynamic_separable_kernels(sig, 0.75 * 9. / np.ones(3), mult)
    for a, b in zip(ks, oks):
        np.testing.assert_array_equal(a, b)
        assert a.shape == (3, 17) and abs(float(a.sum()) - 1.) < 1e-5            # batch-wide normalisation (reference quirk)
    z = mimic_zooms([40, 44, 36], [1., 1., 1.], res[0], [32, 32, 32])
    ds, dz, uz = OG.mimic_acquisition_zooms([40, 44, 36], [1., 1., 1.], res[0], [32, 32, 32])
    np.testing.assert_array_equal(z, np.concatenate([dz, uz, res[0]]).astype(np.float32))


def test_flip_axis_other_than_zero_is_refused():
    """the reference flips along get_ras_axes(aff)[0] (labels_to_image_model.py:159-162); the fused kernel flips axis 0 (what
    BrainGenerator's np.eye(4) selects) -- any other orientation must raise instead of silently flipping the wrong axis."""
    import pytest
    from synthsr_b200.generator import GeneratorPlan
    swapped = np.array([[0, 1., 0, 0], [1., 0, 0, 0], [0, 0, 1., 0], [0, 0, 0, 1.]])
    GeneratorPlan([16, 16, 16], True, 0, GEN_LABELS, None, 1., None, aff=np.eye(4))
    GeneratorPlan([16, 16, 16], True, 0, GEN_LABELS, None, 1., None, aff=swapped, flipping=False)
    with pytest.raises(NotImplementedError, match='flipping along axis 1'):
        GeneratorPlan([16, 16, 16], True, 0, GEN_LABELS, None, 1., None, aff=swapped)
