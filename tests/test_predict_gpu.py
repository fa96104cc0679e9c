"""Inference path on the GPU (SynthSR/predict.py -> ext.neuron.models.unet(input_shape=[None]*3 + [C]) -> tcgen05 forward
with inference-mode BatchNorm) against the torch-CPU oracle.  With the reference's trained weights when the files are
present (baseline/_ref/models, git-ignored copies of /root/reference/models/*.h5 that travel to the GPU box; the test is
skipped where they are absent), with seeded random weights otherwise."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find(rel):
    for base in (os.path.join(ROOT, 'baseline', '_ref'), '/root/reference', '/root/reference/data'):
        p = os.path.join(base, rel)
        if os.path.isfile(p):
            return p
    return None


def _oracle_predict(sd, x, nb_levels):
    from oracle import unet as OU
    params = {k: torch.tensor(np.asarray(v), dtype=torch.float64) for k, v in sd.items()}
    with torch.no_grad():
        return OU.forward(params, torch.from_numpy(x).double(), training=False, nb_levels=nb_levels).numpy()


def test_dynamic_model_predict_matches_oracle_random_weights():
    from ext.neuron import models as nrn_models
    from synthsr_b200.unet import UNet3D
    rng = np.random.default_rng(5)
    for impl, tol in (('ref', 2e-5), ('tc3', 1e-3), ('tc', 4e-3)):     # 'tc' = plain-TF32 fast mode (not the default)
        m = nrn_models.unet(nb_features=8, input_shape=[None, None, None, 2], nb_levels=3, conv_size=3, nb_labels=1, feat_mult=2,
                            nb_conv_per_level=2, final_pred_activation='linear', batch_norm=-1, activation='elu',
                            conv_impl=impl)
        sd = UNet3D([16, 16, 16, 2], nb_features=8, nb_levels=3, seed=4, conv_impl='ref').state_dict()
        for k in sd:                                   # non-trivial BatchNorm parameters / moving statistics
            if k.endswith(('gamma', 'moving_variance')):
                sd[k] = rng.uniform(.5, 1.5, size=sd[k].shape).astype(np.float32)
            if k.endswith(('beta', 'moving_mean')):
                sd[k] = rng.normal(size=sd[k].shape).astype(np.float32) * .3
        m.set_weights(sd)
        for dims in ([32, 48, 32], [16, 32, 64]):     # a second shape rebuilds the engine and keeps the weights
            x = rng.uniform(0, 1, size=(1, *dims, 2)).astype(np.float32)
            pred = m.predict(x)
            ref = _oracle_predict(sd, x, 3)
            err = np.linalg.norm(pred - ref) / np.linalg.norm(ref)
            assert pred.shape == (1, *dims, 1) and err < tol, (impl, dims, err)


def test_real_weights_inference_on_a_real_scan_matches_oracle():
    """models/SynthSR_v10_210712.h5 (13,242,049 trained parameters) on a 64^3 crop of data/images/brain1.nii.gz: the
    tcgen05 forward against the float64 oracle with the same weights; the full predict() pipeline on the whole scan
    produces an in-range 1 mm volume of the input's shape."""
    wfile, image = _find('models/SynthSR_v10_210712.h5'), _find('images/brain1.nii.gz')
    if wfile is None or image is None:
        pytest.skip('reference weights / scan not available on this machine')
    from SynthSR import predict as P
    from ext.lab2im import utils
    from synthsr_b200 import h5lite
    im, aff, _ = utils.load_volume(image, im_only=False, dtype='float')
    S, idx, shape, aff2 = P.preprocess(im, aff)
    model = P.build_unet(1, wfile)
    sd, _ = h5lite.load_keras_weights(wfile)
    c = [s // 2 - 32 for s in S.shape[1:4]]
    crop = np.ascontiguousarray(S[:, c[0]:c[0] + 64, c[1]:c[1] + 64, c[2]:c[2] + 64, :], dtype=np.float32)
    pred = model.predict(crop)
    ref = _oracle_predict(sd, crop, 5)
    err = np.linalg.norm(pred - ref) / np.linalg.norm(ref)
    emax = np.abs(pred - ref).max() / np.abs(ref).max()
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'real_weights_inference.txt'), 'a') as f:
        f.write('real weights, 64^3 crop of brain1: tc3 (default mode) vs float64 oracle rel L2 %.3e, max/max %.3e; pred range [%.3f, %.3f]\n'
                % (err, emax, pred.min(), pred.max()))
    assert err < 1e-3 and emax < 1e-3, (err, emax)      # default mode ('tc3'), north_star bar
    out, aff_out = P.predict_volume(model, im, aff)
    assert out.shape == im.shape and np.isfinite(out).all() and out.min() >= 0 and out.max() <= 128
    assert np.allclose(aff_out, aff)
    # the synthesised MP-RAGE is anatomically aligned with its input: strong correlation inside the head
    mask = im > np.percentile(im, 60)
    r = np.corrcoef(out[mask].ravel(), im[mask].ravel())[0, 1]
    with open(os.path.join(ROOT, 'gpurun_out', 'real_weights_inference.txt'), 'a') as f:
        f.write('full scan %s -> %s, output range [%.1f, %.1f], correlation with the input inside the head %.3f\n'
                % (im.shape, out.shape, out.min(), out.max(), r))
    assert r > 0.3, r
