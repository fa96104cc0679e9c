"""Generates tests/golden/reference_predict_host.npz by EXECUTING the reference's own edit_volumes.resample_volume /
resample_volume_like / align_volume_to_ref (pure NumPy / SciPy; /root/reference, read-only, TF / Keras / nibabel stubbed as
in make_reference_goldens.py) on small seeded volumes with anisotropic, oblique-free affines: the host side of the
inference path (scripts/predict_command_line.py:113-116, predict_command_line_hyperfine.py:110-115).

    python tests/golden/make_reference_predict_goldens.py      # needs /root/reference (build container only)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_reference_goldens import import_reference  # noqa: E402


def cases():
    rng = np.random.default_rng(7)
    out = []
    for shape, vox, flips in [((12, 10, 7), (1.5, 1.5, 5.0), (1, 1, 1)), ((9, 14, 11), (1.0, 1.0, 1.0), (-1, 1, 1)),
                              ((8, 8, 20), (2.0, 0.7, 0.5), (1, -1, 1)), ((10, 6, 9), (1.0, 3.0, 1.2), (1, 1, -1))]:
        vol = rng.uniform(0, 100, size=shape)
        aff = np.eye(4)
        aff[:3, :3] = np.diag(np.array(vox) * np.array(flips))
        aff[:3, 3] = rng.uniform(-20, 20, size=3)
        out.append((vol, aff))
    # a permuted-axes affine (sagittal acquisition) for align_volume_to_ref
    vol = rng.uniform(0, 100, size=(7, 9, 8))
    aff = np.array([[0., 0., 1.2, 3.], [-1., 0., 0., 5.], [0., 2., 0., -4.], [0., 0., 0., 1.]])
    out.append((vol, aff))
    return out


def main():
    _, ev, _, _ = import_reference()
    res = {}
    cs = cases()
    for i, (vol, aff) in enumerate(cs):
        v2, a2 = ev.resample_volume(vol.copy(), aff.copy(), [1.0, 1.0, 1.0])
        v3, a3 = ev.align_volume_to_ref(v2, a2, aff_ref=np.eye(4), return_aff=True, n_dims=3)
        res['vol%d' % i], res['aff%d' % i] = vol, aff
        res['res_vol%d' % i], res['res_aff%d' % i] = v2, a2
        res['ali_vol%d' % i], res['ali_aff%d' % i] = np.ascontiguousarray(v3), a3
    # reslice case 2 (floating) into the aligned grid of case 0 (reference), as the hyperfine script does for the T2
    flo, aff_flo = cs[2]
    like = ev.resample_volume_like(res['ali_vol0'], res['ali_aff0'], flo, aff_flo)
    res['like_0_2'] = like
    np.savez_compressed(os.path.join(HERE, 'reference_predict_host.npz'), **res)
    print({k: v.shape for k, v in res.items() if k.startswith(('res_vol', 'like'))})


if __name__ == '__main__':
    main()
