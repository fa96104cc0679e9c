"""Pins the oracle (and the product's host logic) against golden vectors produced by EXECUTING the reference's own
code in the build container (tests/golden/make_reference_goldens.py: pure-NumPy helpers; make_reference_graph_goldens.py:
the ext/neuron/utils.py graph functions and ext/lab2im gaussian kernels on a NumPy `tf` shim)."""
import json
import os

import numpy as np

from oracle import generator as OG

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, 'golden', 'reference_graph_ops.npz'))
H = json.load(open(os.path.join(HERE, 'golden', 'reference_host_logic.json')))
f32 = np.float32


def _bits(a, b):
    a, b = np.asarray(a, f32), np.asarray(b, f32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_interpn_bit_exact():
    loc = [G['loc'][..., d] for d in range(3)]
    assert _bits(OG.interpn_linear(G['vol'], loc), G['interpn_linear'])
    assert _bits(OG.interpn_nearest(G['vol'], loc), G['interpn_nearest'])


def test_resize_transform_integrate_bit_exact():
    assert _bits(OG.resize(G['small'], [8, 9, 5], 'linear'), G['resize_linear'])
    assert _bits(OG.resize(G['small'], [8, 9, 5], 'nearest'), G['resize_nearest'])
    assert _bits(OG.resize(G['vol'], [3, 4, 2], 'nearest'), G['resize_down_nearest'])
    assert _bits(OG.transform(G['resize_linear'], G['field'], 'linear'), G['transform_linear'])
    assert _bits(OG.integrate_vec(G['field'], 7), G['integrate_vec'])


def test_affine_shift_and_label_warp():
    s = OG.affine_to_shift(G['aff'], [8, 9, 5])
    c = OG.affine_to_shift(G['aff'], [8, 9, 5], G['field'])
    assert _bits(s, G['affine_to_shift']) and _bits(c, G['combine_shift'])   # same pinned left-to-right matmul order
    assert _bits(OG.transform(G['labels'], c, 'nearest'), G['labels_warped'])


def test_gaussian_kernels_bit_exact():
    from synthsr_b200.generator import gaussian_kernel
    assert _bits(OG.gaussian_kernel_dense([.5, .5, .5]), G['gk_05'])
    assert _bits(OG.gaussian_kernel_dense([.42, .42, 1.26], G['gk_mult']), G['gk_acq_jitter'])
    assert _bits(OG.gaussian_kernel_dense([.5, 0., .75]), G['gk_zero_axis'])
    ks = OG.gaussian_kernels_separable([6., .3, 2.1])
    assert _bits(ks[0], G['gk_sep_0']) and ks[1] is None and _bits(ks[2], G['gk_sep_2'])
    # product host code
    assert _bits(gaussian_kernel([.5, .5, .5])[0], G['gk_05'])
    assert _bits(gaussian_kernel([.42, .42, 1.26], G['gk_mult'])[0], G['gk_acq_jitter'])
    assert _bits(gaussian_kernel([.5, 0., .75])[0], G['gk_zero_axis'])
    pk = gaussian_kernel([6., .3, 2.1])
    assert len(pk) == 2 and _bits(pk[0].reshape(-1), G['gk_sep_0']) and _bits(pk[1].reshape(-1), G['gk_sep_2'])


def test_host_logic_matches_reference():
    from ext.lab2im import edit_volumes as EV
    from ext.lab2im import utils as U
    from synthsr_b200 import generator as PG
    for c in H['get_shapes']:
        for fn in (OG.get_shapes, PG.get_shapes):
            crop, out, pad = fn(*c['args'])
            assert list(crop) == c['crop'] and list(out) == c['out'] and (pad if pad is None else list(pad)) == c['pad']
    for c in H['resample_shape']:
        assert OG.get_resample_shape(c['shape'], c['factor']) == c['res'] == U.get_resample_shape(c['shape'], c['factor'])
    for c in H['sigma']:
        for fn in (OG.blurring_sigma_for_downsampling, PG.blurring_sigma):
            np.testing.assert_array_equal(fn(c['cur'], c['down'], c['mult'], c['thick']), c['sigma'])
    for c in H['closest']:
        assert U.find_closest_number_divisible_by_m(c['n'], c['m'], c['t']) == c['res']
    for c in H['ras_axes']:
        assert [int(v) for v in EV.get_ras_axes(np.array(c['aff']))] == c['axes']
    for c in H['align']:
        v2, a2 = EV.align_volume_to_ref(np.array(c['vol']), np.array(c['aff']), aff_ref=np.eye(4), return_aff=True, n_dims=3)
        np.testing.assert_array_equal(v2, np.array(c['aligned']))
        np.testing.assert_allclose(a2, np.array(c['aff_out']))
        np.testing.assert_array_equal(EV.align_volume_to_ref(v2, np.eye(4), aff_ref=np.array(c['aff']), n_dims=3), np.array(c['back']))
    for c in H['padding_margin']:
        assert U.get_padding_margin(c['c'], c['lc']) == c['res']
    for c in H['n_channels_array']:
        np.testing.assert_array_equal(U.reformat_to_n_channels_array(np.array(c['v']) if np.ndim(c['v']) else c['v'], 3, c['nc']), c['res'])
    for c in H['fs_sort']:
        ll, nn = U.get_list_labels(label_list=c['labels'], FS_sort=True)
        assert [int(v) for v in ll] == c['sorted'] and nn == c['n_neutral']
    import tempfile
    for c in H['labels_dir']:                              # folder scan (training() without a label list) + .npz volume info
        d = tempfile.mkdtemp()
        for i, m in enumerate(c['maps']):
            np.savez(os.path.join(d, 'lab%d.npz' % i), vol_data=np.array(m, dtype=np.int32))
        ll, nn = U.get_list_labels(labels_dir=d, FS_sort=c['FS_sort'])
        assert [int(v) for v in ll] == c['labels'] and (None if nn is None else int(nn)) == c['n_neutral']
        info = U.get_volume_info(os.path.join(d, 'lab0.npz'), aff_ref=np.eye(4))
        assert [int(v) for v in info[0]] == c['info_shape'] and int(info[2]) == c['info_n_dims'] and int(info[3]) == c['info_n_channels']
        np.testing.assert_allclose(np.asarray(info[1]), np.array(c['info_aff']))
        np.testing.assert_allclose([float(v) for v in info[5]], c['info_res'])


def test_randomise_res_branch_bit_exact():
    """oracle restatement of the randomise_res branch vs the reference's own edit_tensors.blurring_sigma_for_downsampling
    (tensor branch), gaussian_kernel (sigma as a [B,3] tensor, separable -- including its normalisation by the sum over
    the whole batch) and layers.MimicAcquisition.call, executed on the NumPy tf shim
    (tests/golden/make_reference_randomise_res_goldens.py)."""
    R = np.load(os.path.join(HERE, 'golden', 'reference_randomise_res.npz'))
    sig = OG.dynamic_sigma([1., 1., 1.], R['res'], R['thick'], .42)
    np.testing.assert_array_equal(sig, R['sigma'].astype(np.float32))
    ks = OG.dynamic_separable_kernels(sig, 0.75 * 9. / np.ones(3), R['mult'])
    for i in range(3):
        np.testing.assert_array_equal(ks[i], R['kernel_%d' % i])
    np.testing.assert_array_equal(OG.dynamic_separable_kernels(sig, 0.75 * 9. / np.ones(3), None)[2], R['kernel_nojitter_2'])
    assert abs(float(R['kernel_0'].sum()) - 1.) < 1e-5 and R['kernel_0'].shape == (3, 17)       # one sum for the whole batch
    for b in range(3):
        o, d = OG.mimic_acquisition(R['vol'][b], R['res_mimic'][b], [1., 1., 1.], [1., 1., 1.], [8, 10, 16])
        np.testing.assert_array_equal(o, R['mimic_vol'][b])
        np.testing.assert_array_equal(d, R['mimic_dist'][b])
