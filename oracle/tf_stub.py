"""A numpy-backed stand-in for the handful of TensorFlow 0.x calls that the reference's
``src/ops.py`` makes, so that ``ops.conv2d`` / ``ops.linear`` -- the reference's OWN layer
functions (rows a4-a6 of SURVEY.md §8) -- can be imported and EXECUTED in this container.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Used by oracle/make_golden_network.py.

What this pins and what it does not.  TensorFlow is absent (un-vendored, un-pinned:
README.md:10-14), so the arithmetic INSIDE ``tf.nn.conv2d`` / ``tf.matmul`` /
``tf.nn.bias_add`` is restated here from TF's documented semantics (cross-correlation, filter
``[kh, kw, cin, cout]``, strides ``[1, sh, sw, 1]`` for NHWC, 'VALID' = no padding, float
math; evaluated in float64 so the fixture is not tied to one summation order).  What IS the
reference's executed code is everything ops.py does around those calls: the stride list and
kernel shape it builds from ``x.get_shape()`` (ops.py:14-19), the variable names and shapes it
asks for (ops.py:21-24, 36-39), bias_add after the contraction, the activation applied last
(ops.py:27-28, 43-44) and the ``[in, out]`` orientation of ``linear``'s matrix.  For the head /
loss block of ``src/network.py`` (:60-94) the stub adds placeholder (= its fed value), log,
reduce_sum, one_hot, pow and tensor arithmetic with numpy broadcasting.

Eager: every op computes at once on numpy arrays.  ``get_variable`` hands out the arrays of
``VARIABLES`` (keyed by the variable-scope path, e.g. ``"l1/w"``) and asserts that the shape the
reference requests equals the shape of the array -- the layout contract of the C-ABI's flat
parameter buffer (``arl_param_layout``).
"""
import contextlib
import sys
import types

import numpy as np

VARIABLES = {}        # "scope/name" -> float64 array; filled by the caller
FEEDS = {}            # placeholder name -> array (eager: a placeholder IS its fed value)
REQUESTED = []        # (path, shape) in the order the reference created them
_scope = []


class Shape(object):
    def __init__(self, dims):
        self._d = [int(v) for v in dims]

    def __getitem__(self, i):
        return self._d[i]

    def __len__(self):
        return len(self._d)

    def as_list(self):
        return list(self._d)


class Tensor(object):
    def __init__(self, value):
        self.value = np.asarray(value, np.float64)

    def get_shape(self):
        return Shape(self.value.shape)

    # numpy broadcasting == TF broadcasting ([N] op [N,1] -> [N,N], the reference's D3 hazard)
    def __truediv__(self, other):
        return Tensor(self.value / _val(other))

    __div__ = __truediv__

    def __add__(self, other):
        return Tensor(self.value + _val(other))

    __radd__ = __add__

    def __sub__(self, other):
        return Tensor(self.value - _val(other))

    def __rsub__(self, other):
        return Tensor(_val(other) - self.value)

    def __mul__(self, other):
        return Tensor(self.value * _val(other))

    __rmul__ = __mul__

    def __neg__(self):
        return Tensor(-self.value)


def _val(x):
    return x.value if isinstance(x, Tensor) else np.asarray(x, np.float64)


@contextlib.contextmanager
def variable_scope(name, *a, **k):
    _scope.append(name)
    try:
        yield
    finally:
        _scope.pop()


def get_variable(name, shape=None, dtype=None, initializer=None, **k):
    path = "/".join(_scope + [name])
    shape = [int(v) for v in shape]
    REQUESTED.append((path, tuple(shape)))
    if path not in VARIABLES:
        raise KeyError("the reference asked for variable %r %s; none was provided" % (path, shape))
    v = np.asarray(VARIABLES[path], np.float64)
    assert list(v.shape) == shape, "variable %s: reference wants %s, provided %s" % (
        path, shape, list(v.shape))
    return Tensor(v)


def _conv2d(x, w, strides, padding, data_format="NHWC", **k):
    """tf.nn.conv2d: cross-correlation, filter [kh, kw, cin, cout]."""
    x, w = _val(x), _val(w)
    assert padding == "VALID", padding
    if data_format == "NCHW":
        assert strides[0] == 1 and strides[1] == 1
        x = x.transpose(0, 2, 3, 1)
        sh, sw = strides[2], strides[3]
    else:
        assert data_format == "NHWC" and strides[0] == 1 and strides[3] == 1
        sh, sw = strides[1], strides[2]
    n, h, wd, c = x.shape
    kh, kw, cin, cout = w.shape
    assert cin == c
    oh, ow = (h - kh) // sh + 1, (wd - kw) // sw + 1
    out = np.zeros((n, oh, ow, cout), np.float64)
    for i in range(kh):
        for j in range(kw):
            patch = x[:, i:i + sh * (oh - 1) + 1:sh, j:j + sw * (ow - 1) + 1:sw, :]   # [n, oh, ow, c]
            out += patch @ w[i, j]
    if data_format == "NCHW":
        out = out.transpose(0, 3, 1, 2)
    return Tensor(out)


def _bias_add(x, b, data_format=None, **k):
    x, b = _val(x), _val(b)
    if data_format == "NCHW" and x.ndim == 4:
        return Tensor(x + b[None, :, None, None])
    return Tensor(x + b)                       # last axis (NHWC, and every 2-D input)


def _relu(x, **k):
    return Tensor(np.maximum(_val(x), 0.0))


def _matmul(a, b, **k):
    return Tensor(_val(a) @ _val(b))


def _reshape(x, shape, **k):
    return Tensor(_val(x).reshape([int(v) for v in shape]))


def _softmax(x, **k):
    v = _val(x)
    e = np.exp(v - v.max(axis=-1, keepdims=True))
    return Tensor(e / e.sum(axis=-1, keepdims=True))


def _placeholder(dtype, shape=None, name=None):
    if name not in FEEDS:
        raise KeyError("placeholder %r has no fed value" % (name,))
    v = np.asarray(FEEDS[name], np.float64)
    if shape is not None:
        assert v.ndim == len(shape) and all(d is None or int(d) == n for d, n in zip(shape, v.shape)), (
            name, shape, v.shape)
    return Tensor(v)


def _log(x, **k):
    return Tensor(np.log(_val(x)))


def _reduce_sum(x, reduction_indices=None, keep_dims=False, **k):
    return Tensor(_val(x).sum(axis=reduction_indices, keepdims=keep_dims))


def _one_hot(indices, depth, on_value=1.0, off_value=0.0, **k):
    idx = np.asarray(_val(indices)).astype(np.int64)
    out = np.full(idx.shape + (int(depth),), float(off_value))
    np.put_along_axis(out, idx[..., None], float(on_value), axis=-1)
    return Tensor(out)


def _pow(x, y, **k):
    return Tensor(_val(x) ** _val(y))


def _square(x, **k):
    return Tensor(_val(x) ** 2)


def _reduce_mean(x, reduction_indices=None, keep_dims=False, **k):
    return Tensor(_val(x).mean(axis=reduction_indices, keepdims=keep_dims))


def _initializer(*a, **k):
    return ("initializer", a, tuple(sorted(k.items())))


def install():
    """Register the stub as ``tensorflow`` (and the sub-modules ops.py imports) in sys.modules."""
    tf = types.ModuleType("tensorflow")
    tf.float32 = "float32"
    tf.variable_scope = variable_scope
    tf.get_variable = get_variable
    tf.matmul = _matmul
    tf.reshape = _reshape
    tf.placeholder = _placeholder
    tf.log = _log
    tf.reduce_sum = _reduce_sum
    tf.one_hot = _one_hot
    tf.pow = _pow
    tf.square = _square
    tf.reduce_mean = _reduce_mean
    tf.constant_initializer = _initializer
    tf.random_normal_initializer = _initializer
    tf.truncated_normal_initializer = _initializer
    nn = types.ModuleType("tensorflow.nn")
    nn.conv2d, nn.bias_add, nn.relu, nn.softmax = _conv2d, _bias_add, _relu, _softmax
    tf.nn = nn
    contrib = types.ModuleType("tensorflow.contrib")
    layers = types.ModuleType("tensorflow.contrib.layers")
    layers.xavier_initializer = _initializer
    python = types.ModuleType("tensorflow.contrib.layers.python")
    pylayers = types.ModuleType("tensorflow.contrib.layers.python.layers")
    pylayers.initializers = types.ModuleType("tensorflow.contrib.layers.python.layers.initializers")
    tf.contrib, contrib.layers, layers.python, python.layers = contrib, layers, python, pylayers
    for m in (tf, nn, contrib, layers, python, pylayers, pylayers.initializers):
        sys.modules[m.__name__] = m
    return tf
