"""Host side of the inference path (SynthSR/predict.py, ext/lab2im/edit_volumes.py) against outputs of the reference's own
functions executed in the build container (tests/golden/make_reference_predict_goldens.py)."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, 'golden', 'reference_predict_host.npz'))


@pytest.mark.parametrize('i', range(5))
def test_resample_and_align_match_reference(i):
    from ext.lab2im import edit_volumes
    v2, a2 = edit_volumes.resample_volume(G['vol%d' % i].copy(), G['aff%d' % i].copy(), [1.0, 1.0, 1.0])
    assert v2.shape == G['res_vol%d' % i].shape
    np.testing.assert_allclose(v2, G['res_vol%d' % i], rtol=0, atol=1e-10)
    np.testing.assert_allclose(a2, G['res_aff%d' % i], rtol=0, atol=1e-12)
    v3, a3 = edit_volumes.align_volume_to_ref(v2, a2, aff_ref=np.eye(4), return_aff=True, n_dims=3)
    np.testing.assert_array_equal(np.ascontiguousarray(v3), G['ali_vol%d' % i])
    np.testing.assert_allclose(a3, G['ali_aff%d' % i], rtol=0, atol=1e-12)


def test_resample_volume_like_matches_reference():
    from ext.lab2im import edit_volumes
    like = edit_volumes.resample_volume_like(G['ali_vol0'], G['ali_aff0'], G['vol2'], G['aff2'])
    np.testing.assert_allclose(like, G['like_0_2'], rtol=0, atol=1e-10)


def test_pre_and_post_processing_follow_the_script():
    """predict_command_line.py:110-135: normalisation to [0, 1], centred zero padding to multiples of 32, 255 x, clip to
    [0, 128], crop back."""
    from SynthSR import predict as P
    rng = np.random.default_rng(3)
    im = rng.uniform(-50, 900, size=(37, 50, 33))
    aff = np.diag([1., 1., 1., 1.])
    S, idx, shape, aff2 = P.preprocess(im, aff)
    assert S.shape == (1, 64, 64, 64, 1) and list(idx) == [13, 7, 15] and shape == (1, 37, 50, 33, 1)
    core = S[0, 13:50, 7:57, 15:48, 0]
    assert core.min() == 0.0 and core.max() == 1.0 and S.sum() == core.sum()           # zero padding around the scan
    out = np.zeros(S.shape)
    out[0, 13:50, 7:57, 15:48, 0] = np.linspace(-0.2, 1.0, core.size).reshape(core.shape)
    pred = P.postprocess(out, idx, shape)
    assert pred.shape == (37, 50, 33) and pred.min() == 0.0 and pred.max() == 128.0
    S_ct, *_ = P.preprocess(np.array(im), aff, ct=True)                                   # CT window [0, 80] first
    assert S_ct.max() == 1.0
    with pytest.raises(Exception, match='extension not supported'):
        open('/tmp/_ssr_bad.txt', 'w').write('x')
        P._io_lists('/tmp/_ssr_bad.txt', '/tmp/out')


def test_dynamic_unet_model_checks_shapes_and_keeps_weights_by_name(tmp_path):
    from ext.neuron import models as nrn_models
    from synthsr_b200 import h5lite
    from synthsr_b200.unet import layer_specs
    m = nrn_models.unet(nb_features=4, input_shape=[None, None, None, 1], nb_levels=2, conv_size=3, nb_labels=1,
                        feat_mult=2, nb_conv_per_level=2, final_pred_activation='linear', batch_norm=-1, activation='elu')
    assert isinstance(m, nrn_models.DynamicUnetModel) and m.layer_names[0] == 'unet_conv_downarm_0_0'
    rng = np.random.default_rng(0)
    w = {}
    for name, kind, ci, co in layer_specs(1, 4, 2):
        if kind == 'bn':
            for s in ('gamma', 'beta', 'moving_mean', 'moving_variance'):
                w['%s/%s' % (name, s)] = rng.uniform(.5, 1.5, size=co).astype(np.float32)
        else:
            k = 3 if kind == 'conv' else 1
            w[name + '/kernel'] = rng.normal(size=(k, k, k, ci, co)).astype(np.float32)
            w[name + '/bias'] = rng.normal(size=co).astype(np.float32)
    p = str(tmp_path / 'w.h5')
    h5lite.save_keras_weights(p, w)
    m.load_weights(p, by_name=True)
    assert all(np.array_equal(m.get_weights()[k], w[k]) for k in w)
    w['unet_conv_downarm_0_0/kernel'] = np.zeros((3, 3, 3, 2, 4), np.float32)            # a 2-channel model's first layer
    h5lite.save_keras_weights(p, w)
    with pytest.raises(ValueError, match='not compatible'):
        m.load_weights(p, by_name=True)


class _StandInUnet:
    """the deterministic stand-in network of tests/golden/make_reference_predict_script_goldens.py (same arithmetic)."""

    def load_weights(self, path, by_name=True):
        self.loaded = path

    def predict(self, S):
        S = np.asarray(S, dtype=np.float64)
        g = [np.arange(n, dtype=np.float64) for n in S.shape[1:4]]
        ramp = (np.sin(.37 * g[0])[:, None, None] + .5 * np.cos(.21 * g[1])[None, :, None] + .002 * g[2][None, None, :] ** 1.5)
        out = .55 * S[..., 0] + .25 * np.roll(S[..., 0], 2, axis=1) * (1 + .1 * ramp) - .03 + .02 * ramp
        if S.shape[-1] > 1:
            out = out * .3 - .2 * S[..., 1] + .1 * np.roll(S[..., 1], 1, axis=2)
        return out[..., None]


def test_predict_glue_matches_the_reference_scripts_run_end_to_end(monkeypatch):
    """the reference's own scripts/predict_command_line.py (default + --ct --disable_flipping) and
    scripts/predict_command_line_hyperfine.py executed with runpy around a stand-in network and in-memory volume I/O
    (tests/golden/make_reference_predict_script_goldens.py) vs SynthSR.predict with the same stand-in: saved volume and saved
    affine.  Pure float64 NumPy on both sides -> 1e-9."""
    import SynthSR.predict as P
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_predict_scripts.npz'))
    net = _StandInUnet()
    pred, aff = P.predict_volume(net, G['a_im'], G['a_aff'])
    assert pred.shape == G['a_pred'].shape
    np.testing.assert_allclose(pred, G['a_pred'], rtol=0, atol=1e-9)
    np.testing.assert_allclose(aff, G['a_pred_aff'], rtol=0, atol=1e-9)
    pred, aff = P.predict_volume(net, G['b_im'], G['b_aff'], ct=True, disable_flipping=True)
    np.testing.assert_allclose(pred, G['b_pred'], rtol=0, atol=1e-9)
    np.testing.assert_allclose(aff, G['b_pred_aff'], rtol=0, atol=1e-9)
    # flip test-time augmentation really matters for this stand-in (else the comparison above would not see it)
    p2, _ = P.predict_volume(net, G['a_im'], G['a_aff'], disable_flipping=True)
    assert np.abs(p2 - G['a_pred']).max() > 1.

    vols = {'/t1.nii.gz': (G['h_t1'], G['h_t1_aff']), '/t2.nii.gz': (G['h_t2'], G['h_t2_aff'])}
    saved = {}
    monkeypatch.setattr(P, 'build_unet', lambda *a, **k: net)
    monkeypatch.setattr(P.utils, 'load_volume', lambda path, im_only=True, dtype=None, **kw: (
        vols[path][0].astype(np.float64).copy(), vols[path][1].copy(), None))
    monkeypatch.setattr(P.utils, 'save_volume', lambda vol, aff, hdr, path, **kw: saved.__setitem__(path, (np.array(vol), np.array(aff))))
    real_isfile = os.path.isfile
    monkeypatch.setattr(os.path, 'isfile', lambda p: p in vols or real_isfile(p))
    P.predict_hyperfine('/t1.nii.gz', '/t2.nii.gz', '/pred_h.nii.gz')
    pred, aff = saved['/pred_h.nii.gz']
    np.testing.assert_allclose(pred, G['h_pred'], rtol=0, atol=1e-9)
    np.testing.assert_allclose(aff, G['h_pred_aff'], rtol=0, atol=1e-9)
