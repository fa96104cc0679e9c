"""The oracle's intensity / label stages against outputs of the reference's OWN Keras layers (ext/lab2im/layers.py)
executed in the build container on the NumPy `tf` shim with injected random draws
(tests/golden/make_reference_layer_goldens.py -> tests/golden/reference_layers.npz)."""
import os

import numpy as np

from oracle import generator as OG

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, 'golden', 'reference_layers.npz'))
GEN = np.array([0, 14, 15, 16, 2, 3, 4, 5, 7, 8, 10, 11, 12, 13, 17, 18, 26, 28, 31])
f32 = np.float32


def test_sample_conditional_gmm_bit_exact_including_the_batch_sum():
    """SampleConditionalGMM (layers.py:480-498), two channels.  Batch 2 exposes what tf.scatter_nd does to the tiled
    indices: the look-up table is the SUM of the parameters over the batch (kept by the oracle and by the product's
    generator.gmm_luts)."""
    from synthsr_b200.generator import gmm_luts
    for B in (1, 2):
        lab, m, s, n, ref = (G['gmm%d_%s' % (B, k)] for k in ('labels', 'means', 'stds', 'noise', 'out'))
        for b in range(B):
            for i in range(2):
                got = OG.sample_conditional_gmm(lab[b, ..., 0], m[:, :, i], s[:, :, i], n[b, ..., i], GEN)
                np.testing.assert_array_equal(got, ref[b, ..., i])
        for i in range(2):                                    # the product's host-side tables give the same image
            ml, sl = gmm_luts(m[:, :, i], s[:, :, i], GEN, int(GEN.max()) + 1)
            for b in range(B):
                got = (sl[b][lab[b, ..., 0]] * n[b, ..., i]).astype(f32) + ml[b][lab[b, ..., 0]]
                np.testing.assert_array_equal(got.astype(f32), ref[b, ..., i])
    # batch 2 really differs from per-example tables
    per_b = OG.sample_conditional_gmm(G['gmm2_labels'][0, ..., 0], G['gmm2_means'][:1, :, 0], G['gmm2_stds'][:1, :, 0],
                                      G['gmm2_noise'][0, ..., 0], GEN)
    assert np.abs(per_b - G['gmm2_out'][0, ..., 0]).max() > 1.


def test_bias_field_corruption_bit_exact():
    """BiasFieldCorruption(.3, .2, False) (layers.py:1067-1097): applied (u < .95) and skipped."""
    for tag in ('on', 'off'):
        draws = {'bias_std_0': G['bias_%s_std' % tag].reshape(1), 'bias_normal_0': G['bias_%s_normal' % tag][..., 0],
                 'bias_apply_0': tag == 'on'}
        assert list(G['bias_%s_normal' % tag].shape[1:4]) == [int(v) for v in G['bias_small_shape'][:3]]
        got = OG.bias_field_corruption(G['bias_%s_x' % tag][0, ..., 0], draws, 0, 0, .3, .2)
        np.testing.assert_array_equal(got, G['bias_%s_out' % tag][0, ..., 0])
    assert np.array_equal(G['bias_off_out'], G['bias_off_x'])


def test_intensity_augmentation_bit_exact():
    """IntensityAugmentation(clip=300, normalise=True, gamma_std=.5, separate_channels=True) (layers.py:1186-1257)."""
    for b in range(2):
        gamma = f32(f32(G['int_gamma'][b].reshape(-1)[0]) * f32(.5))
        got = OG.intensity_augmentation(G['int_x'][b, ..., 0], clip=300, gamma=gamma)
        np.testing.assert_array_equal(got, G['int_out'][b, ..., 0])


def test_random_flip_and_swap_bit_exact():
    """RandomFlip(0, [True, False], label_list, n_neutral) (layers.py:391-427): flip draws .2 / .7 / .49 against prob .5."""
    labels, image, u = G['flip_labels'], G['flip_image'], G['flip_u']
    ll = G['flip_label_list']
    for b in range(3):
        flip = bool(u[b, 0] < .5)
        got_l = OG.random_flip(labels[b, ..., 0], flip, ll, 3, swap=True)
        got_i = OG.random_flip(image[b, ..., 0], flip, ll, 3, swap=False)
        np.testing.assert_array_equal(got_l, G['flip_out_labels'][b, ..., 0])
        np.testing.assert_array_equal(got_i, G['flip_out_image'][b, ..., 0])
    assert not np.array_equal(G['flip_out_labels'][0], labels[0][::-1])       # sided labels were swapped, not only mirrored


def test_random_crop_bit_exact_and_draw_convention():
    """RandomCrop (layers.py:252-270): offsets = int32(uniform(0, in - crop)) (truncation), the same for all inputs --
    which is how synthsr_b200.draws.sample_draws forms 'crop_idx'."""
    a, b_, u, cs = G['crop_a'], G['crop_b'], G['crop_u'], [int(v) for v in G['crop_shape']]
    for b in range(2):
        idx = u[b].astype(f32).astype(np.int32)
        np.testing.assert_array_equal(OG.random_crop(a[b, ..., 0], idx, cs), G['crop_out_a'][b, ..., 0])
        np.testing.assert_array_equal(OG.random_crop(b_[b, ..., 0], idx, cs), G['crop_out_b'][b, ..., 0])


def test_gaussian_blur_matches_reference_layer():
    """GaussianBlur(.5) and GaussianBlur(.42 * [1, 1, 3], 1.15) (layers.py:732-767; tf.nn.conv3d 'SAME' = zero padding).
    The summation order of tf.nn.conv3d is not specified: a few ulp."""
    x = G['blur_x'][0, ..., 0]
    np.testing.assert_allclose(OG.gaussian_blur(x, .5), G['blur_out_05'][0, ..., 0], rtol=0, atol=5e-7)
    got = OG.gaussian_blur(x, list(G['blur_sigma']), G['blur_mult'])
    np.testing.assert_allclose(got, G['blur_out_acq'][0, ..., 0], rtol=0, atol=5e-7)


S = np.load(os.path.join(HERE, 'golden', 'reference_spatial.npz'))


def test_sample_affine_transform_bit_exact():
    """ext/lab2im/utils.py sample_affine_transform / create_rotation_transform / create_shearing_transform executed on the
    shim with the training() bounds (rotation, shearing, scaling, translation draws injected): the oracle's and the
    product's affine builders give the same 4x4 matrices bit for bit."""
    from synthsr_b200 import draws as D
    for b in range(3):
        args = (S['aff_rotation'][b], S['aff_shearing'][b], S['aff_scaling'][b], S['aff_translation'][b])
        np.testing.assert_array_equal(OG.build_affine(*args), S['aff_out'][b])
        np.testing.assert_array_equal(D.build_affine(*args), S['aff_out'][b])


def test_random_spatial_deformation_layer_bit_exact():
    """the complete RandomSpatialDeformation layer of the reference (layers.py:161-211: affine sampling, SVF draw -> Resize ->
    VecInt -> Resize, SpatialTransformer with the non-linear field first) on labels ('nearest') and an image ('linear'),
    plus its elastic-only and affine-only configurations."""
    shape = list(S['rsd_labels'].shape[1:4])
    lab_in = S['rsd_labels'][0].astype(f32)
    assert OG.get_resample_shape(shape, .0625) == [int(v) for v in S['rsd_small_shape'][:3]]
    draws = {'svf_std': S['rsd_svf_std'].reshape(-1)[0], 'svf_normal': S['rsd_svf_normal']}
    field, _ = OG.random_spatial_deformation_field(draws, 0, shape, .0625)
    aff = OG.build_affine(S['rsd_rotation'][0], S['rsd_shearing'][0], S['rsd_scaling'][0], S['rsd_translation'][0])
    lab = OG.spatial_transformer(lab_in, aff, field, 'nearest')[..., 0].astype(np.int32)
    img = OG.spatial_transformer(S['rsd_image'][0], aff, field, 'linear')[..., 0]
    np.testing.assert_array_equal(lab, S['rsd_out_labels'][0, ..., 0])
    np.testing.assert_array_equal(img, S['rsd_out_image'][0, ..., 0])
    assert (lab != S['rsd_labels'][0, ..., 0]).mean() > .5                      # the deformation really moved the labels
    draws = {'svf_std': S['el_svf_std'].reshape(-1)[0], 'svf_normal': S['el_svf_normal']}
    field, _ = OG.random_spatial_deformation_field(draws, 0, shape, .0625)
    lab = OG.transform(lab_in, field, 'nearest')[..., 0].astype(np.int32)
    np.testing.assert_array_equal(lab, S['el_out_labels'][0, ..., 0])
    aff = OG.build_affine(S['af_rotation'][0], S['af_shearing'][0], S['af_scaling'][0], None)
    lab = OG.spatial_transformer(lab_in, aff, None, 'nearest')[..., 0].astype(np.int32)
    np.testing.assert_array_equal(lab, S['af_out_labels'][0, ..., 0])
