"""Extension of tf_numpy_shim for the reference's Keras LAYERS (ext/lab2im/layers.py): the ops their `call` methods use on
top of the spatial-transform subset -- scatter_nd, slice, reverse, switch / less, reductions, floormod, conv3d -- again as
the IEEE float32 NumPy equivalents, with every tf.random draw popped from an injected queue (uniform: the value itself,
which must already lie in [minval, maxval); normal: a standard normal scaled by stddev and shifted by mean in float32, the
way the TF kernel applies them)."""
import types

import numpy as np

import tf_numpy_shim as base


def install(random_queue):
    tf, K, T = base.install(random_queue)
    _np, _t = base._np, base._t

    def split(x, sizes, axis=0):
        x = _np(x)
        if isinstance(sizes, (int, np.integer)):
            return [_t(a) for a in np.split(x, sizes, axis=axis)]
        sizes = [int(s) for s in sizes]
        if -1 in sizes:
            sizes[sizes.index(-1)] = x.shape[axis] - (sum(sizes) + 1)
        return [_t(a) for a in np.split(x, np.cumsum(sizes)[:-1], axis=axis)]

    def scatter_nd(indices, updates, shape):
        out = np.zeros([int(s) for s in _np(shape).reshape(-1)], dtype=_np(updates).dtype)
        idx = _np(indices)
        np.add.at(out, tuple(idx[..., d] for d in range(idx.shape[-1])), _np(updates))      # duplicates ADD, like TF
        return _t(out)

    def slice_(x, begin, size):
        x = _np(x)
        b = [int(v) for v in _np(begin).reshape(-1)]
        s = [int(v) for v in _np(size).reshape(-1)]
        return _t(x[tuple(slice(bi, None if si == -1 else bi + si) for bi, si in zip(b, s))])

    def where(c, a=None, b=None):
        if a is None:
            return _t(np.argwhere(_np(c)).astype(np.int64))
        return _t(np.where(_np(c), _np(a), _np(b)))

    def switch(cond, a, b):
        c = _np(cond)
        assert c.size == 1, 'K.switch is only used with scalar conditions in the reference layers'
        return a if bool(c.reshape(-1)[0]) else b

    def uniform(shape, minval=0, maxval=1, dtype='float32'):
        v = _np(random_queue.pop(0)).astype(np.float32)
        want = tuple(int(s) for s in _np(shape).reshape(-1))
        assert v.shape == want, ('uniform draw shape', v.shape, want)
        lo, hi = np.asarray(_np(minval), np.float32), np.asarray(_np(maxval), np.float32)
        assert np.all(v >= lo) and np.all(v <= hi), 'injected uniform draw outside [minval, maxval]'
        return _t(v)

    def normal(shape, mean=0., stddev=1., dtype='float32'):
        v = _np(random_queue.pop(0)).astype(np.float32)
        want = tuple(int(s) for s in _np(shape).reshape(-1))
        assert v.shape == want, ('normal draw shape', v.shape, want)
        return _t((v * np.asarray(_np(stddev), np.float32) + np.asarray(_np(mean), np.float32)).astype(np.float32))

    def conv3d(x, k, strides=None, padding='SAME'):
        """NDHWC x [kd,kh,kw,Cin,Cout], zero 'SAME' padding, float32 accumulation in tap order (TF's own order is not
        specified: compared with a few-ulp tolerance)."""
        x, k = _np(x).astype(np.float32), _np(k).astype(np.float32)
        kd, kh, kw, ci, co = k.shape
        p = [(s - 1) // 2 for s in (kd, kh, kw)]
        xp = np.pad(x, ((0, 0), (p[0], kd - 1 - p[0]), (p[1], kh - 1 - p[1]), (p[2], kw - 1 - p[2]), (0, 0)))
        out = np.zeros(x.shape[:4] + (co,), np.float32)
        for a in range(kd):
            for b in range(kh):
                for c in range(kw):
                    patch = xp[:, a:a + x.shape[1], b:b + x.shape[2], c:c + x.shape[3], :]
                    out = (out + np.einsum('bxyzi,io->bxyzo', patch, k[a, b, c]).astype(np.float32)).astype(np.float32)
        return _t(out)

    import sys
    Layer = sys.modules['keras.layers'].Layer

    def layer_call(self, inputs, **kwargs):            # Keras Layer.__call__: build on first use from the input shapes
        if not getattr(self, 'built', False):
            def shape_of(v):                             # Keras hands build() the STATIC shape (batch None in a graph)
                return tuple(v.get_shape()) if isinstance(v, T) else tuple(_np(v).shape)
            shp = [shape_of(v) for v in inputs] if isinstance(inputs, (list, tuple)) else shape_of(inputs)
            self.build(shp)
        return self.call(inputs, **kwargs)

    Layer.__call__ = layer_call
    def matmul(a, b):
        """2-D as in the base shim; batched [B,n,k] @ [B,k,m] with the same pinned left-to-right float32 accumulation."""
        a, b = _np(a), _np(b)
        if a.ndim == 2:
            return base._matmul(a, b)
        return _t(np.stack([_np(base._matmul(a[i], b[i])) for i in range(a.shape[0])]))

    def diag(x):
        x = _np(x)
        out = np.zeros(x.shape + (x.shape[-1],), x.dtype)
        for i in range(x.shape[-1]):
            out[..., i, i] = x[..., i]
        return _t(out)

    tf.matmul = matmul
    tf.linalg = types.SimpleNamespace(diag=diag)
    tf.eye = lambda n, dtype='float32': _t(np.eye(int(n), dtype=np.float32))
    tf.split = split
    tf.scatter_nd = scatter_nd
    tf.slice = slice_
    tf.where = where
    tf.squeeze = lambda x, axis=None: _t(np.squeeze(_np(x), axis=axis))
    tf.reverse = lambda x, axis: _t(np.flip(_np(x), axis=tuple(int(a) for a in _np(axis).reshape(-1))))
    tf.less = lambda a, b: _t(_np(a) < _np(b))
    tf.sort = lambda x, axis=-1: _t(np.sort(_np(x), axis=axis))
    tf.cos = lambda x: _t(np.cos(_np(x)))
    tf.sin = lambda x: _t(np.sin(_np(x)))
    tf.abs = lambda x: _t(np.abs(_np(x)))
    tf.maximum = lambda a, b: _t(np.maximum(_np(a), _np(b)))
    tf.minimum = lambda a, b: _t(np.minimum(_np(a), _np(b)))
    tf.reduce_max = lambda x, axis=None, keepdims=False: _t(np.max(_np(x), axis=axis, keepdims=keepdims))
    tf.reduce_min = lambda x, axis=None, keepdims=False: _t(np.min(_np(x), axis=axis, keepdims=keepdims))
    tf.logical_and = lambda a, b: _t(np.logical_and(_np(a), _np(b)))
    tf.logical_not = lambda a: _t(np.logical_not(_np(a)))
    tf.math.multiply = lambda a, b: _t((_np(a) * _np(b)).astype(np.result_type(_np(a).dtype, np.float32)
                                                             if np.issubdtype(_np(a).dtype, np.floating) else _np(a) * _np(b)))
    tf.math.floormod = lambda a, b: _t(np.mod(_np(a), b))
    tf.math.reduce_max = tf.reduce_max
    tf.math.reduce_min = tf.reduce_min
    tf.math.maximum = tf.maximum
    tf.nn = types.SimpleNamespace(conv3d=conv3d)
    tf.random = types.SimpleNamespace(uniform=uniform, normal=normal)
    K.less = lambda a, b: _t(_np(a) < np.asarray(_np(b), _np(a).dtype))
    K.greater = lambda a, b: _t(_np(a) > np.asarray(_np(b), _np(a).dtype))
    K.switch = switch
    K.min = lambda x, axis=None: _t(np.min(_np(x), axis=tuple(axis) if isinstance(axis, list) else axis))
    K.max = lambda x, axis=None: _t(np.max(_np(x), axis=tuple(axis) if isinstance(axis, list) else axis))
    K.tile = lambda x, n: tf.tile(x, n)
    K.concatenate = lambda xs, axis=-1: tf.concat(xs, axis)
    K.cast = lambda x, dtype: tf.cast(x, dtype)
    K.flatten = lambda x: _t(_np(x).reshape(-1))
    K.exp = tf.exp
    K.zeros_like = tf.zeros_like
    K.ones_like = tf.ones_like
    K.int_shape = lambda x: tuple(_np(x).shape)
    return tf, K, T
