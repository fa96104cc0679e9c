"""A tiny NumPy-backed stand-in for the handful of TensorFlow / Keras-backend ops that the reference's spatial-transform
code (ext/neuron/utils.py, ext/lab2im/edit_tensors.gaussian_kernel) calls, so that the UNMODIFIED reference functions can
be executed in this container (TensorFlow 2.0 is not installable) to produce golden vectors for the oracle.

Semantics: every op is the IEEE float32 element-wise NumPy equivalent of the TF kernel (tf.round = half-to-even,
tf.clip_by_value = min/max, tf.gather = fancy indexing, tf.cast float->int truncates).  The only op whose TF
summation order is not reproducible is tf.matmul (4x4 @ 4xV in affine_to_shift); it is evaluated with the oracle's
pinned left-to-right float32 order and that golden is compared with a tolerance of a few ulp.
"""
import sys
import types

import numpy as np


class TensorShape(tuple):
    def as_list(self):
        return list(self)

    def __getitem__(self, item):
        r = tuple.__getitem__(self, item)
        return TensorShape(r) if isinstance(item, slice) else r


GRAPH_BATCH = [None]      # set by make_reference_model_goldens.py while the reference builds its graph


class T(np.ndarray):
    """ndarray that answers the TF tensor API used by the reference."""

    def __new__(cls, a, dtype=None):
        return np.asarray(a, dtype=dtype).view(cls)

    def get_shape(self):
        # `static_batch_unknown` mimics a Keras tensor whose batch dimension is None at graph-construction time
        shp = np.ndarray.shape.__get__(self)
        if getattr(self, 'static_batch_unknown', False):
            return TensorShape((None,) + tuple(shp[1:]))
        if GRAPH_BATCH[0] is not None and len(shp) >= 2 and shp[0] == GRAPH_BATCH[0]:
            return TensorShape((None,) + tuple(shp[1:]))       # whole-graph runs: every batched tensor has batch dim None
        return TensorShape(shp)

    @property
    def shape(self):
        return TensorShape(np.ndarray.shape.__get__(self))

    @property
    def dtype(self):
        return _DT(np.ndarray.dtype.__get__(self))

    # TF tensors are immutable: `a *= b` rebinds the name to a NEW tensor (prod_n in ext/neuron/utils.py relies on it)
    def __iadd__(self, o): return self + o
    def __isub__(self, o): return self - o
    def __imul__(self, o): return self * o
    def __itruediv__(self, o): return self / o

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        """TF converts python / NumPy operands to the tensor's dtype (no float64 promotion)."""
        fdt = None
        for x in inputs:
            if isinstance(x, T) and np.issubdtype(np.ndarray.dtype.__get__(x), np.floating):
                fdt = np.ndarray.dtype.__get__(x)
        conv = []
        for x in inputs:
            if isinstance(x, T):
                conv.append(x.view(np.ndarray))
            elif fdt is not None and np.issubdtype(np.asarray(x).dtype, np.floating):
                conv.append(np.asarray(x, dtype=fdt))
            elif fdt is not None and np.issubdtype(np.asarray(x).dtype, np.integer) and not isinstance(x, np.ndarray):
                conv.append(np.asarray(x, dtype=fdt))
            else:
                conv.append(np.asarray(x))
        if out is not None:
            kwargs['out'] = tuple(o.view(np.ndarray) if isinstance(o, T) else o for o in out)
        r = getattr(ufunc, method)(*conv, **kwargs)
        if isinstance(r, tuple):
            return tuple(T(v) if isinstance(v, np.ndarray) else v for v in r)
        return T(r) if isinstance(r, np.ndarray) else (T(r) if np.ndim(r) == 0 and method == '__call__' else r)


class _DT:
    def __init__(self, d):
        self.d = np.dtype(d)

    def __eq__(self, o):
        return self.d == np.dtype(o.d if isinstance(o, _DT) else o)

    def __ne__(self, o):
        return not self.__eq__(o)

    def __hash__(self):
        return hash(self.d)

    @property
    def base_dtype(self):
        return self


def _np(x):
    return np.asarray(x)


def _t(x, dtype=None):
    return T(np.asarray(x, dtype=dtype))


def _dtype(d):
    return d.d if isinstance(d, _DT) else np.dtype(d)


def _matmul(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    out = np.zeros((a.shape[0], b.shape[1]), np.float32)
    for i in range(a.shape[0]):
        acc = (a[i, 0] * b[0]).astype(np.float32)
        for k in range(1, a.shape[1]):
            acc = (acc + (a[i, k] * b[k]).astype(np.float32)).astype(np.float32)
        out[i] = acc
    return _t(out)


def install(random_queue=None):
    """registers fake `tensorflow`, `keras`, `keras.backend`, ... modules in sys.modules."""
    tf = types.ModuleType('tensorflow')
    tf.__path__ = []
    tf.TensorShape = TensorShape
    tf.float32, tf.int32, tf.float64, tf.bool = 'float32', 'int32', 'float64', 'bool'
    tf.is_tensor = lambda x: isinstance(x, T)
    tf.cast = lambda x, dtype: _t(np.trunc(_np(x)) if np.issubdtype(_dtype(dtype), np.integer) and
                                  np.issubdtype(_np(x).dtype, np.floating) else _np(x)).astype(_dtype(dtype)).view(T)
    tf.floor = lambda x: _t(np.floor(_np(x)))
    tf.round = lambda x: _t(np.rint(_np(x)))
    tf.clip_by_value = lambda x, lo, hi: _t(np.minimum(np.maximum(_np(x), np.asarray(lo, _np(x).dtype)),
                                                       np.asarray(hi, _np(x).dtype)))
    tf.stack = lambda xs, axis=0: _t(np.stack([_np(x) for x in xs], axis=axis))
    tf.unstack = lambda x, axis=0: [_t(np.take(_np(x), i, axis=axis)) for i in range(_np(x).shape[axis])]
    tf.reshape = lambda x, shape: _t(np.reshape(_np(x), [int(s) for s in np.asarray(shape).reshape(-1)]))
    tf.transpose = lambda x, perm=None: _t(np.transpose(_np(x), perm))
    tf.gather = lambda p, idx, axis=0: _t(np.take(_np(p), _np(idx), axis=axis))
    tf.range = lambda a, b=None: _t(np.arange(a, b, dtype=np.int32) if b is not None else np.arange(a, dtype=np.int32))
    tf.tile = lambda x, m: _t(np.tile(_np(x), [int(v) for v in np.asarray(m).reshape(-1)]))
    tf.size = lambda x: int(_np(x).size)
    tf.shape = lambda x: _t(np.array(_np(x).shape, np.int32))
    tf.ones = lambda shape, dtype='float32': _t(np.ones([int(s) for s in np.asarray(shape).reshape(-1)], _dtype(dtype)))
    tf.zeros = lambda shape, dtype='float32': _t(np.zeros([int(s) for s in np.asarray(shape).reshape(-1)], _dtype(dtype)))
    tf.ones_like = lambda x: _t(np.ones_like(_np(x)))
    tf.zeros_like = lambda x: _t(np.zeros_like(_np(x)))
    tf.matmul = _matmul
    tf.concat = lambda xs, axis: _t(np.concatenate([_np(x) for x in xs], axis=axis))
    tf.split = lambda x, sizes, axis=0: [_t(a) for a in np.split(_np(x), np.cumsum(sizes)[:-1], axis=axis)]
    tf.expand_dims = lambda x, axis: _t(np.expand_dims(_np(x), axis))
    tf.convert_to_tensor = lambda x, dtype=None: _t(x, None if dtype is None else _dtype(dtype))
    tf.exp = lambda x: _t(np.exp(_np(x)))
    tf.equal = lambda a, b: _t(_np(a) == _np(b))
    tf.where = lambda c, a=None, b=None: _t(np.where(_np(c), _np(a), _np(b)))
    tf.reduce_sum = lambda x, axis=None, keepdims=False: _t(np.sum(_np(x), axis=axis, keepdims=keepdims, dtype=_np(x).dtype))
    def map_fn(fn, elems, dtype=None):
        saved, GRAPH_BATCH[0] = GRAPH_BATCH[0], None           # inside the mapped function the tensors have no batch dimension
        try:
            it = zip(*elems) if isinstance(elems, (list, tuple)) else elems
            res = [fn(list(e)) if isinstance(e, tuple) else fn(e) for e in it]
        finally:
            GRAPH_BATCH[0] = saved
        if isinstance(res[0], (list, tuple)):
            return [_t(np.stack([_np(r[j]) for r in res])) for j in range(len(res[0]))]
        return _t(np.stack([_np(r) for r in res]))

    tf.map_fn = map_fn
    tf.math = types.SimpleNamespace(log=lambda x: _t(np.log(_np(x))), exp=tf.exp, minimum=lambda a, b: _t(np.minimum(_np(a), _np(b))),
                                    equal=tf.equal, pow=lambda a, b: _t(np.power(_np(a), _np(b))),
                                    floor=lambda x: _t(np.floor(_np(x))), ceil=lambda x: _t(np.ceil(_np(x))),
                                    sqrt=lambda x: _t(np.sqrt(_np(x))), square=lambda x: _t(np.square(_np(x))),
                                    reduce_sum=tf.reduce_sum)
    q = random_queue if random_queue is not None else []
    tf.random = types.SimpleNamespace(
        uniform=lambda shape, minval=0, maxval=1, dtype='float32': _t(q.pop(0)),
        normal=lambda shape, mean=0., stddev=1.: _t((_np(q.pop(0)) * np.float32(stddev) + np.float32(mean)).astype(np.float32)))
    K = types.ModuleType('keras.backend')
    K.expand_dims = lambda x, axis=-1: _t(np.expand_dims(_np(x), axis))
    K.square = lambda x: _t(np.square(_np(x)))
    K.sum = lambda x, axis=None: _t(np.sum(_np(x), axis=axis, dtype=_np(x).dtype))
    K.reshape = lambda x, s: tf.reshape(x, s)
    K.permute_dimensions = lambda x, p: _t(np.transpose(_np(x), p))
    K.epsilon = lambda: 1e-7
    K.clip = lambda x, lo, hi: _t(np.minimum(np.maximum(_np(x), np.asarray(_np(lo), _np(x).dtype)), np.asarray(_np(hi), _np(x).dtype)))

    class _Any(types.ModuleType):
        __path__ = []

        def __getattr__(self, item):
            if item.startswith('__'):
                raise AttributeError(item)
            return type(item, (object,), {})

    sys.modules['tensorflow'] = tf
    sys.modules['keras.backend'] = K
    keras = _Any('keras')
    keras.backend = K
    sys.modules['keras'] = keras
    for name in ['keras.layers', 'keras.models', 'keras.engine', 'keras.engine.topology', 'nibabel', 'keras.callbacks',
                 'keras.optimizers', 'keras.initializers', 'tensorflow.keras', 'keras.legacy', 'keras.legacy.interfaces',
                 'keras.utils', 'keras.constraints', 'keras.regularizers']:
        sys.modules[name] = _Any(name)
    keras.layers = sys.modules['keras.layers']

    class Lambda:                                      # KL.Lambda(fn)(x) -> fn(x)
        def __init__(self, fn, **kwargs):
            self.fn = fn

        def __call__(self, x):
            return self.fn(x)

    class Layer:                                       # base class of the reference's custom layers
        def __init__(self, **kwargs):
            pass

        def build(self, input_shape):
            self.built = True

    keras.layers.Lambda = Lambda
    keras.layers.Layer = Layer
    sys.modules['keras.engine'].Layer = Layer
    sys.modules['keras.engine.topology'].Layer = Layer
    np.int, np.float = int, float
    import scipy.stats
    if not hasattr(scipy.stats, 'median_absolute_deviation'):
        scipy.stats.median_absolute_deviation = scipy.stats.median_abs_deviation
    return tf, K, T
