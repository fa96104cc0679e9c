"""Generates tests/golden/reference_host_logic.json by EXECUTING the reference's own pure-NumPy helpers
(/root/reference, read-only) in this container.  TensorFlow / Keras / nibabel are not installable here, so they are
replaced by empty stub modules: only functions that never touch them are called (shape bookkeeping, blurring sigmas,
orientation helpers, label sorting, padding margins).  The graph ops themselves (tf.*) cannot be executed -> the
float kernels stay "parity unpinned"; this file pins the host-side logic the oracle and the product both restate.

    python tests/golden/make_reference_goldens.py          # needs /root/reference (build container only)
"""
import json
import os
import sys
import types

import numpy as np

REF = '/root/reference'


class _Stub(types.ModuleType):
    """module whose every missing attribute is a dummy class (enough for `from keras.x import Y` at import time)."""
    __path__ = []

    def __getattr__(self, item):
        if item.startswith('__'):
            raise AttributeError(item)
        return type(item, (object,), {})


def _stub(name):
    m = _Stub(name)
    sys.modules[name] = m
    return m


def import_reference():
    np.int, np.float = int, float                      # aliases removed in NumPy 2, used by the reference
    for name in ['tensorflow', 'keras', 'keras.layers', 'keras.backend', 'keras.models', 'keras.engine',
                 'keras.engine.topology', 'nibabel', 'keras.callbacks', 'keras.optimizers', 'keras.initializers',
                 'tensorflow.keras', 'keras.legacy', 'keras.legacy.interfaces', 'keras.utils', 'keras.constraints',
                 'keras.regularizers']:
        _stub(name)
    sys.modules['keras.layers'].Layer = object
    sys.modules['keras.engine'].Layer = object
    sys.modules['keras.engine.topology'].Layer = object
    sys.modules['keras.engine'].InputSpec = object
    sys.modules['keras.models'].Model = object
    sys.modules['keras'].layers = sys.modules['keras.layers']
    sys.modules['keras'].backend = sys.modules['keras.backend']
    sys.modules['tensorflow'].is_tensor = lambda x: False
    import scipy.stats
    if not hasattr(scipy.stats, 'median_absolute_deviation'):
        scipy.stats.median_absolute_deviation = scipy.stats.median_abs_deviation      # removed upstream
    sys.path.insert(0, REF)
    from ext.lab2im import utils, edit_volumes, edit_tensors
    from SynthSR import labels_to_image_model as l2i
    return utils, edit_volumes, edit_tensors, l2i


def main():
    utils, ev, et, l2i = import_reference()
    out = {'get_shapes': [], 'resample_shape': [], 'sigma': [], 'closest': [], 'ras_axes': [], 'align': [],
           'padding_margin': [], 'n_channels_array': [], 'fs_sort': []}
    for args in [([148, 187, 155], None, [1., 1., 1.], [1., 1., 1.], None, 32),
                 ([148, 187, 155], 128, [1., 1., 1.], [1., 1., 1.], None, 32),
                 ([160, 160, 160], None, [1., 1., 1.], [1., 1., 1.], None, 32),
                 ([192, 192, 64], None, [1., 1., 1.], [1., 1., 1.], None, 32),
                 ([148, 187, 155], [96, 128, 96], [1., 1., 1.], [2., 2., 2.], None, 32),
                 ([40, 48, 36], 32, [1., 1., 1.], [1., 1., 1.], 4, None),
                 ([100, 100, 100], None, [1., 1., 1.], [.5, .5, .5], None, 16),
                 ([100, 90, 80], 64, [1., 1., 1.], [1.5, 1.5, 1.5], 2, 8)]:
        crop, outs, pad = l2i.get_shapes(*args)
        out['get_shapes'].append({'args': args, 'crop': [int(v) for v in crop], 'out': [int(v) for v in outs],
                                  'pad': None if pad is None else [int(v) for v in pad]})
    # randomised sweep of the same function (seeded): odd shapes, anisotropic atlas / target resolutions, scalar and list crops
    rng = np.random.default_rng(2024)
    res_choices = [.5, .7, 1., 1.2, 1.5, 2., 3.]
    for _ in range(200):
        shape = [int(v) for v in rng.integers(17, 200, size=3)]
        atlas = [float(rng.choice(res_choices))] * 3 if rng.uniform() < .6 else [float(v) for v in rng.choice(res_choices, 3)]
        target = list(atlas) if rng.uniform() < .4 else ([float(rng.choice(res_choices))] * 3 if rng.uniform() < .6 else
                                                         [float(v) for v in rng.choice(res_choices, 3)])
        u = rng.uniform()
        oshape = None if u < .3 else (int(rng.integers(16, 160)) if u < .65 else [int(v) for v in rng.integers(16, 160, size=3)])
        pad = None if rng.uniform() < .6 else (int(rng.integers(1, 9)) if rng.uniform() < .5 else
                                               [int(v) for v in rng.integers(0, 9, size=3)])
        div = None if rng.uniform() < .3 else int(rng.choice([2, 4, 8, 16, 32]))
        args = (shape, oshape, atlas, target, pad, div)
        crop, outs, pd = l2i.get_shapes(*args)
        out['get_shapes'].append({'args': list(args), 'crop': [int(v) for v in crop], 'out': [int(v) for v in outs],
                                  'pad': None if pd is None else [int(v) for v in pd]})
    for shape, f in [([148, 187, 155], .03125), ([148, 187, 155], .0625), ([160] * 3, .03125), ([160] * 3, .025),
                     ([256] * 3, .025), ([192, 192, 64], .025), ([64] * 3, .0625), ([128] * 3, [.03125, .0625, .025])]:
        out['resample_shape'].append({'shape': shape, 'factor': f, 'res': utils.get_resample_shape(shape, f)})
    for cur, down, mult, thick in [([1., 1., 1.], [1., 1., 1.], None, None), ([1., 1., 1.], [1., 1., 1.], .42, [1., 1., 1.]),
                                   ([1., 1., 1.], [1., 1., 3.], .42, [1., 1., 3.]), ([1., 1., 1.], [1.5, 1.5, 5.], .42, [1.5, 1.5, 5.]),
                                   ([1., 1., 1.], [1., 1., 6.], .42, [1., 1., 4.]), ([1., 1., 1.], [2., 2., 2.], None, None),
                                   ([1., 1., 1.], [1., 2., 0.], None, None)]:
        s = et.blurring_sigma_for_downsampling(cur, down, mult, thick)
        out['sigma'].append({'cur': cur, 'down': down, 'mult': mult, 'thick': thick, 'sigma': [float(v) for v in s]})
    for n, m, t in [(187, 32, 'lower'), (160, 32, 'lower'), (155, 32, 'lower'), (100, 16, 'closer'), (100, 16, 'higher')]:
        out['closest'].append({'n': n, 'm': m, 't': t, 'res': int(utils.find_closest_number_divisible_by_m(n, m, t))})
    affs = [np.eye(4), np.array([[0, 0, -1, 10], [1, 0, 0, 5], [0, -1, 0, 3], [0, 0, 0, 1.]]),
            np.array([[-1, 0, 0, 90], [0, 0, 1, -20], [0, -1, 0, 30], [0, 0, 0, 1.]]),
            np.array([[.9, .1, 0, 0], [-.1, .9, .05, 0], [0, -.05, 1.1, 0], [0, 0, 0, 1.]])]
    rng = np.random.default_rng(0)
    vol = rng.integers(0, 50, size=(4, 5, 6)).astype(np.int32)
    for a in affs:
        out['ras_axes'].append({'aff': a.tolist(), 'axes': [int(v) for v in ev.get_ras_axes(a)]})
        v2, a2 = ev.align_volume_to_ref(vol, a, aff_ref=np.eye(4), return_aff=True, n_dims=3)
        v3 = ev.align_volume_to_ref(v2, np.eye(4), aff_ref=a, n_dims=3)
        out['align'].append({'aff': a.tolist(), 'vol': vol.tolist(), 'aligned': np.asarray(v2).tolist(),
                             'aff_out': np.asarray(a2).tolist(), 'back': np.asarray(v3).tolist()})
    for c, lc in [(160, 128), ([160, 128, 96], 64), (None, 16), (96, None)]:
        out['padding_margin'].append({'c': c, 'lc': lc, 'res': utils.get_padding_margin(c, lc)})
    for v, nc in [(1., 1), ([1., 1., 3.], 2), (np.array([[1.5, 1.5, 5.], [1., 1., 1.]]), 2), (np.array([1.0004, 2., 3.]), 1)]:
        r = utils.reformat_to_n_channels_array(v, 3, nc)
        out['n_channels_array'].append({'v': np.asarray(v).tolist(), 'nc': nc, 'res': np.asarray(r).tolist()})
    for labels in [[0, 14, 15, 16, 2, 3, 4, 41, 42, 43], [0, 2, 3, 4, 5, 7, 8, 10, 11, 12, 13, 14, 15, 16, 17, 18, 26, 28, 31],
                   [0, 24, 4, 43, 17, 53, 2, 41]]:
        ll, nn = utils.get_list_labels(label_list=labels, FS_sort=True)
        out['fs_sort'].append({'labels': labels, 'sorted': [int(v) for v in ll], 'n_neutral': int(nn)})
    # get_list_labels scanning a folder of label maps (training() without path_generation_labels) and get_volume_info on .npz
    import tempfile
    out['labels_dir'] = []
    rng = np.random.default_rng(5)
    for pools, fs in [([[0, 2, 3, 41, 42], [0, 3, 4, 14, 43], [0, 16, 24, 2, 41]], True),
                      ([[0, 5, 9, 1], [7, 3, 0]], False)]:
        d = tempfile.mkdtemp()
        maps = []
        for i, pool in enumerate(pools):
            m = np.array(pool)[rng.integers(0, len(pool), size=(4, 5, 3))].astype(np.int32)
            m.reshape(-1)[:len(pool)] = pool
            np.savez(os.path.join(d, 'lab%d.npz' % i), vol_data=m)
            maps.append(m.tolist())
        ll, nn = utils.get_list_labels(labels_dir=d, FS_sort=fs)
        info = utils.get_volume_info(os.path.join(d, 'lab0.npz'), aff_ref=np.eye(4))
        out['labels_dir'].append({'maps': maps, 'FS_sort': fs, 'labels': [int(v) for v in ll], 'n_neutral': None if nn is None else int(nn),
                                  'info_shape': [int(v) for v in info[0]], 'info_aff': np.asarray(info[1]).tolist(),
                                  'info_n_dims': int(info[2]), 'info_n_channels': int(info[3]),
                                  'info_res': [float(v) for v in info[5]]})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_host_logic.json')
    json.dump(out, open(path, 'w'), indent=1)
    print('wrote', path, {k: len(v) for k, v in out.items()})


if __name__ == '__main__':
    main()
