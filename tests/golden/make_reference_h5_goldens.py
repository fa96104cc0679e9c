"""Pins synthsr_b200/h5lite.py against the reference's own Keras weight files (models/SynthSR_v10_210712*.h5, written by
h5py 2.10 / Keras 2.3.1): per-tensor shape + float64 sum + sum of squares + first/last raw values, the layer_names order
and the root attributes.  Run in the build container (needs /root/reference); the summary travels as a small fixture.

    python tests/golden/make_reference_h5_goldens.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from synthsr_b200 import h5lite  # noqa: E402

REF = '/root/reference/models'


def summarise(path):
    w, attrs = h5lite.load_keras_weights(path)
    f = h5lite.H5File(path)
    out = {'file_bytes': os.path.getsize(path),
           'layer_names': [n.decode() for n in attrs['layer_names']],
           'backend': attrs['backend'].decode(), 'keras_version': attrs['keras_version'].decode(),
           'n_tensors': len(w), 'n_params': int(sum(v.size for v in w.values())), 'tensors': {}}
    for k, v in w.items():
        v64 = v.astype(np.float64).ravel()
        out['tensors'][k] = {'shape': list(v.shape), 'dtype': str(v.dtype), 'sum': float(v64.sum()),
                             'sumsq': float((v64 * v64).sum()), 'first': float(v64[0]), 'last': float(v64[-1])}
    # raw weight_names of two layers (TensorFlow-uniquified variable scopes in the hyperfine file)
    out['weight_names'] = {l: [n.decode() for n in np.atleast_1d(f[l].attrs['weight_names'])]
                           for l in ('unet_conv_downarm_0_0', 'unet_bn_down_1')}
    return out


if __name__ == '__main__':
    res = {fn: summarise(os.path.join(REF, fn)) for fn in sorted(os.listdir(REF)) if fn.endswith('.h5')}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_h5_summary.json'), 'w') as fh:
        json.dump(res, fh, indent=1)
    for fn, r in res.items():
        print(fn, r['n_tensors'], r['n_params'])
