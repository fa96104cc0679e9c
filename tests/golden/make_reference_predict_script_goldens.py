"""Golden vectors for the inference glue: the reference's OWN command-line scripts (scripts/predict_command_line.py and
scripts/predict_command_line_hyperfine.py) are run with runpy on anisotropic, obliquely-oriented scans.  Replaced around
them, because they cannot exist here: `tensorflow` (only used for a thread setting), the Keras U-Net (a deterministic
stand-in model whose output depends on position and is NOT left/right symmetric, so padding offsets, the flip test-time
augmentation and the residual arithmetic all show), and the nibabel file I/O of utils.load_volume / save_volume (volumes +
affines are handed over in memory).  Everything between load and save is the reference's code, unmodified: CT clipping,
resample_volume to 1 mm, align_volume_to_ref, normalisation, centred zero-padding to multiples of 32, flip averaging,
rescaling / clipping, cropping, and the Hyperfine script's T2 resampling, intensity scalings and residual.

Note: predict_command_line.py reads `args.model` / `args.disable_flipping` from the dict returned by vars() (:79, :126), which
raises AttributeError as shipped; the harness hands argparse a namespace whose __dict__ also answers attribute access, so
the script runs as written.

Writes tests/golden/reference_predict_scripts.npz.   (build container only: needs /root/reference)"""
import argparse
import os
import runpy
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_numpy_shim  # noqa: E402

tf, K, T = tf_numpy_shim.install([])
import types  # noqa: E402

tf.config = types.SimpleNamespace(threading=types.SimpleNamespace(set_intra_op_parallelism_threads=lambda n: None))
sys.path.insert(0, '/root/reference')
from ext.lab2im import utils  # noqa: E402
from ext.neuron import models as nrn_models  # noqa: E402


class StandInUnet:
    """deterministic 'network': position-dependent, asymmetric along axis 1 (the flipped axis), mixes both channels."""

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


nrn_models.unet = lambda **kw: StandInUnet()


class AttrDict(dict):
    __getattr__ = dict.__getitem__


_parse = argparse.ArgumentParser.parse_args


def parse_args(self, *a, **k):
    ns = _parse(self, *a, **k)
    ns.__dict__ = AttrDict(ns.__dict__)
    return ns


argparse.ArgumentParser.parse_args = parse_args

VOLUMES, SAVED = {}, {}
utils.load_volume = lambda path, im_only=True, dtype=None, **kw: (VOLUMES[path][0].astype(np.float64).copy(),
                                                                  VOLUMES[path][1].copy(), None)
utils.save_volume = lambda vol, aff, hdr, path, **kw: SAVED.__setitem__(path, (np.array(vol), np.array(aff)))
_isfile = os.path.isfile
os.path.isfile = lambda p: p in VOLUMES or _isfile(p)

rng = np.random.default_rng(41)


def scan(shape, vox, perm_flip):
    """smooth blobby volume with an affine: voxel sizes `vox`, axes permuted / flipped, plus a translation."""
    g = np.meshgrid(*[np.linspace(0, 1, s) for s in shape], indexing='ij')
    vol = 60 + 40 * np.sin(5 * g[0] + 1) * np.cos(4 * g[1]) + 30 * g[2] + rng.normal(size=shape) * 3
    aff = np.zeros((4, 4))
    perm, flip = perm_flip
    for i in range(3):
        aff[perm[i], i] = vox[i] * flip[i]
    aff[:3, 3] = rng.uniform(-40, 40, size=3)
    aff[3, 3] = 1
    return vol, aff


out = {}
# ---- predict_command_line.py: (a) default, flip TTA; (b) --ct --disable_flipping --model X -----------------------------
VOLUMES['/scan_a.nii.gz'] = scan((20, 26, 9), (1.3, 1.1, 4.0), ((0, 1, 2), (1, 1, 1)))
VOLUMES['/scan_b.nii.gz'] = scan((24, 10, 22), (1.0, 3.5, 1.2), ((2, 0, 1), (-1, 1, -1)))
VOLUMES['/scan_b.nii.gz'] = (VOLUMES['/scan_b.nii.gz'][0] * 2 - 60, VOLUMES['/scan_b.nii.gz'][1])     # values outside [0, 80]
for tag, argv in (('a', ['/scan_a.nii.gz', '/pred_a.nii.gz']),
                  ('b', ['/scan_b.nii.gz', '/pred_b.nii.gz', '--ct', '--disable_flipping', '--model', '/some/model.h5'])):
    sys.argv = ['/root/reference/scripts/predict_command_line.py'] + argv
    runpy.run_path(sys.argv[0], run_name='__main__')
    out['%s_im' % tag], out['%s_aff' % tag] = VOLUMES[argv[0]]
    out['%s_pred' % tag], out['%s_pred_aff' % tag] = SAVED[argv[1]]
# ---- predict_command_line_hyperfine.py ------------------------------------------------------------------------------------
VOLUMES['/t1.nii.gz'] = scan((22, 24, 8), (1.5, 1.5, 5.0), ((0, 1, 2), (-1, 1, 1)))
VOLUMES['/t2.nii.gz'] = scan((18, 20, 10), (1.6, 1.7, 4.0), ((0, 1, 2), (-1, 1, 1)))
VOLUMES['/t2.nii.gz'][1][:3, 3] = VOLUMES['/t1.nii.gz'][1][:3, 3] + [1.5, -2., 3.]        # registered: same physical region
sys.argv = ['/root/reference/scripts/predict_command_line_hyperfine.py', '/t1.nii.gz', '/t2.nii.gz', '/pred_h.nii.gz']
runpy.run_path(sys.argv[0], run_name='__main__')
out['h_t1'], out['h_t1_aff'] = VOLUMES['/t1.nii.gz']
out['h_t2'], out['h_t2_aff'] = VOLUMES['/t2.nii.gz']
out['h_pred'], out['h_pred_aff'] = SAVED['/pred_h.nii.gz']
for k, v in out.items():
    print(k, np.asarray(v).shape)
np.savez_compressed(os.path.join(HERE, 'reference_predict_scripts.npz'), **out)
