"""synthsr_b200/h5lite.py: pure-Python HDF5 subset for Keras weight files.

Pinned against the reference's own files (models/SynthSR_v10_210712*.h5, written by h5py 2.10 / Keras 2.3.1): the
committed summary tests/golden/reference_h5_summary.json (made by tests/golden/make_reference_h5_goldens.py) is checked
against the architecture the engine builds, and -- where /root/reference is mounted -- against a live read."""
import json
import os

import numpy as np
import pytest

from synthsr_b200 import h5lite
from synthsr_b200.unet import keras_layer_order, layer_specs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_h5_summary.json')
REF_MODELS = '/root/reference/models'


def _expected_shapes(cin):
    shapes = {}
    for name, kind, ci, co in layer_specs(cin):
        if kind == 'bn':
            for w in ('gamma', 'beta', 'moving_mean', 'moving_variance'):
                shapes['%s/%s' % (name, w)] = [co]
        else:
            k = 3 if kind == 'conv' else 1
            shapes[name + '/kernel'] = [k, k, k, ci, co]
            shapes[name + '/bias'] = [co]
    return shapes


@pytest.mark.parametrize('fn,cin,nparams', [('SynthSR_v10_210712.h5', 1, 13242049),
                                           ('SynthSR_v10_210712_hyperfine.h5', 2, 13242697)])
def test_reference_files_match_engine_architecture(fn, cin, nparams):
    """tensor names / shapes / count of the shipped Keras models == the engine's parameter layout (SURVEY.md U6:
    13,242,049 parameters for Cin = 1), layer_names == keras_layer_order()."""
    g = json.load(open(GOLD))[fn]
    assert g['n_params'] == nparams and g['keras_version'] == '2.3.1' and g['backend'] == 'tensorflow'
    assert g['layer_names'] == keras_layer_order(5, 2)
    exp = _expected_shapes(cin)
    assert {k: v['shape'] for k, v in g['tensors'].items()} == exp
    assert list(g['tensors'].keys()) == list(exp.keys())            # same order as the engine's flat buffer + moving stats
    assert all(v['dtype'] == 'float32' for v in g['tensors'].values())
    for k, v in g['tensors'].items():                                 # plausible trained values
        n = int(np.prod(v['shape']))
        if k.endswith('moving_variance'):
            assert v['sum'] > 0
        if k.endswith('kernel'):
            assert 0 < v['sumsq'] / n < 1.0


@pytest.mark.skipif(not os.path.isdir(REF_MODELS), reason='reference models are only mounted in the build container')
@pytest.mark.parametrize('fn', ['SynthSR_v10_210712.h5', 'SynthSR_v10_210712_hyperfine.h5'])
def test_live_read_of_reference_files_and_rewrite(fn, tmp_path):
    g = json.load(open(GOLD))[fn]
    path = os.path.join(REF_MODELS, fn)
    w, attrs = h5lite.load_keras_weights(path)
    assert [n.decode() for n in attrs['layer_names']] == g['layer_names']
    for k, v in w.items():
        t = g['tensors'][k]
        v64 = v.astype(np.float64).ravel()
        assert list(v.shape) == t['shape'] and v64.sum() == t['sum'] and (v64 * v64).sum() == t['sumsq']
        assert v64[0] == t['first'] and v64[-1] == t['last']
    f = h5lite.H5File(path)
    assert [n.decode() for n in f['unet_bn_down_1'].attrs['weight_names']] == g['weight_names']['unet_bn_down_1']
    # re-written by our writer: identical content, and byte-identical header messages for a dataset
    out = str(tmp_path / 'rewrite.h5')
    h5lite.save_keras_weights(out, w, layer_order=g['layer_names'])
    w2, a2 = h5lite.load_keras_weights(out)
    assert list(w2.keys()) == list(w.keys()) and all(np.array_equal(w[k], w2[k]) for k in w)
    f2 = h5lite.H5File(out)
    ds = 'unet_conv_downarm_0_0/unet_conv_downarm_0_0/bias:0'
    ds_ref = 'unet_conv_downarm_0_0/' + g['weight_names']['unet_conv_downarm_0_0'][1]   # TF may have uniquified the scope
    ours = {t: bytes(b) for t, b in f2[ds]._msgs}
    ref = {t: bytes(b) for t, b in f[ds_ref]._msgs}
    for t in (0x0001, 0x0003, 0x0005):                  # dataspace, datatype, fill value: byte-identical to h5py's
        assert ours[t] == ref[t], hex(t)
    assert ours[0x0008][:2] == ref[0x0008][:2] and ours[0x0008][10:18] == ref[0x0008][10:18]   # contiguous layout, size
    wn_ours = [b for t, b in f2['unet_conv_downarm_0_0']._msgs if t == 0x000C][0]
    wn_ref = [b for t, b in f['unet_conv_downarm_0_0']._msgs if t == 0x000C][0]
    if fn == 'SynthSR_v10_210712.h5':
        assert bytes(wn_ours) == bytes(wn_ref)          # the weight_names attribute message, byte for byte
    assert open(out, 'rb').read(24) == open(path, 'rb').read(24)      # superblock prefix (versions, sizes, B-tree K)


def test_generic_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    root = h5lite.Group(attrs={'title': 'abc', 'n': np.int64(7), 'vec': np.arange(5, dtype=np.float64),
                               'names': np.array([b'a', b'bcd', b'ef'])})
    g1 = h5lite.Group(attrs={'empty': np.zeros((0,), np.float64)})
    for i in range(23):                                  # > 8 links: several symbol-table nodes under one B-tree
        g1.children['d%02d' % i] = rng.normal(size=(i + 1, 3)).astype(np.float32)
    root.children['many'] = g1
    root.children['nested'] = h5lite.Group(children={'deeper': h5lite.Group(children={
        'i32': np.arange(12, dtype=np.int32).reshape(3, 4), 'f64': rng.normal(size=(2, 2, 2)), 'u8': np.arange(5, dtype=np.uint8),
        'scalar': np.float32(3.5)})})
    root.children['hollow'] = h5lite.Group()
    p = str(tmp_path / 't.h5')
    h5lite.write_h5(p, root)
    f = h5lite.H5File(p)
    assert sorted(f.keys()) == ['hollow', 'many', 'nested'] and f['hollow'].keys() == []
    assert f.attrs['title'] == b'abc' and f.attrs['n'] == 7 and np.array_equal(f.attrs['vec'], np.arange(5.))
    assert list(f.attrs['names']) == [b'a', b'bcd', b'ef']
    assert f['many'].attrs['empty'].shape == (0,)
    assert f['many'].keys() == ['d%02d' % i for i in range(23)]
    for i in range(23):
        assert np.array_equal(f['many/d%02d' % i][()], g1.children['d%02d' % i])
    d = root.children['nested'].children['deeper'].children
    for k in ('i32', 'f64', 'u8'):
        got = f['nested/deeper/' + k]
        assert got.dtype == d[k].dtype and got.shape == d[k].shape and np.array_equal(got[()], d[k])
    assert f['nested/deeper/scalar'][()] == np.float32(3.5)
    with pytest.raises(KeyError):
        f['nested/nope']


def test_keras_layouts_and_optimizer_state(tmp_path):
    rng = np.random.default_rng(1)
    w = {}
    for name, kind, ci, co in layer_specs(1, nb_features=4, nb_levels=2):
        if kind == 'bn':
            for s in ('gamma', 'beta', 'moving_mean', 'moving_variance'):
                w['%s/%s' % (name, s)] = rng.normal(size=co).astype(np.float32)
        else:
            k = 3 if kind == 'conv' else 1
            w[name + '/kernel'] = rng.normal(size=(k, k, k, ci, co)).astype(np.float32)
            w[name + '/bias'] = rng.normal(size=co).astype(np.float32)
    order = keras_layer_order(2)
    for full in (False, True):
        p = str(tmp_path / ('w%d.h5' % full))
        extra = {'m': rng.normal(size=10).astype(np.float32), 'iterations': np.array([42], dtype=np.int64)}
        h5lite.save_keras_weights(p, w, order, extra=extra, full_model=full)
        w2, attrs = h5lite.load_keras_weights(p)
        assert list(w2) == list(w) and all(np.array_equal(w[k], w2[k]) for k in w)
        assert [n.decode() for n in attrs['layer_names']] == order
        f = h5lite.H5File(p)
        assert ('model_weights' in f.keys()) == full
        g = f['model_weights'] if full else f
        assert g['unet_maxpool_0'].keys() == [] and g['unet_maxpool_0'].attrs['weight_names'].shape == (0,)
        assert list(g['unet_bn_down_0'].attrs['weight_names']) == [
            b'unet_bn_down_0/gamma:0', b'unet_bn_down_0/beta:0', b'unet_bn_down_0/moving_mean:0',
            b'unet_bn_down_0/moving_variance:0']
        ex = h5lite.load_extra(p)
        assert int(ex['iterations'][0]) == 42 and np.array_equal(ex['m'], extra['m'])


def test_rejects_what_it_does_not_implement(tmp_path):
    p = str(tmp_path / 'bad.h5')
    open(p, 'wb').write(b'not an hdf5 file at all' * 10)
    with pytest.raises(h5lite.H5FormatError):
        h5lite.H5File(p)
    data = bytearray(h5lite._Writer().finish(h5lite.Group()))
    data[8] = 2                                          # superblock version 2 (new-style files)
    open(p, 'wb').write(bytes(data))
    with pytest.raises(h5lite.H5FormatError):
        h5lite.H5File(p)
