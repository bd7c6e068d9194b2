"""numpy stand-in for the dozen TensorFlow-1.x ops the reference's generation graph calls
(reference modules.py:11-60,129-270; models.py:23-78,105-136) -- TEST INFRASTRUCTURE ONLY.

It exists so that the reference's OWN `modules.py` / `models.py` can be imported and executed in
this container (TensorFlow 1.x cannot be installed) to produce golden vectors: the model wiring,
variable names and op order then come from the reference's source, only the op kernels below are
restated. Each op evaluates eagerly on numpy arrays; `tf.get_variable` resolves names through the
active `variable_scope` stack against a weight dict installed with `set_variables`.
"""
import contextlib
import types

import numpy as np

float32 = np.float32
AUTO_REUSE = object()


class _T(np.ndarray):
    """ndarray with the one Tensor method the reference calls on activations (modules.py:275: input.get_shape())."""

    def get_shape(self):
        return tuple(int(s) for s in self.shape)


def _t(x):
    return np.asarray(x).view(_T)

_scope = []
_variables = {}
_created = []          # names in creation order (the reference's variable list)
_logistic_sample = [None]


def set_variables(mapping):
    _variables.clear()
    _variables.update(mapping)
    del _created[:]


def created_variables():
    return list(_created)


def set_logistic_sample(array):
    _logistic_sample[0] = array


@contextlib.contextmanager
def variable_scope(name, reuse=None, **_):
    _scope.append(name)
    try:
        yield
    finally:
        _scope.pop()


name_scope = variable_scope


def zeros_initializer(*a, **k):
    return 'zeros'


def ones_initializer(*a, **k):
    return 'ones'


def get_variable(name, shape=None, initializer=None, trainable=True, **_):
    full = '/'.join(_scope + [name])
    if full not in _variables:
        raise KeyError('tf_shim: no value for variable ' + full)
    value = np.asarray(_variables[full])
    if shape is not None and tuple(int(s) for s in shape) != value.shape:
        raise ValueError('tf_shim: %s has shape %s, graph asks for %s' % (full, value.shape, tuple(shape)))
    if full not in _created:
        _created.append(full)
    return _t(value)


def trainable_variables(scope=None):
    return [n for n in _created if scope is None or n.startswith(scope)]


def add_to_collection(*a, **k):
    return None


class GraphKeys(object):
    UPDATE_OPS = 'update_ops'


def shape(x):
    return tuple(int(s) for s in np.shape(x))


def div(a, b):
    return a // b


def pad(value, paddings, **_):
    return _t(np.pad(value, [tuple(int(v) for v in p) for p in paddings]))


def reshape(x, shape, **_):
    return _t(np.reshape(x, [int(s) for s in shape]))


def transpose(x, perm=None, **_):
    return _t(np.transpose(x, perm))


def slice(x, begin, size, **_):     # noqa: A001 (mirrors tf.slice)
    idx = []
    for b, s, dim in zip(begin, size, np.shape(x)):
        idx.append(np.s_[int(b):(dim if int(s) == -1 else int(b) + int(s))])
    return _t(x[tuple(idx)])


def tile(x, multiples, **_):
    return _t(np.tile(x, [int(m) for m in multiples]))


def expand_dims(x, axis, **_):
    return _t(np.expand_dims(x, axis))


def squeeze(x, axis=None, **_):
    return _t(np.squeeze(x, axis))


def tanh(x, **_):
    return _t(np.tanh(x))


def sigmoid(x, **_):
    return _t(1.0 / (1.0 + np.exp(-x)))


def add(a, b, **_):
    return _t(a + b)


def _conv1d(value, filters, stride=1, padding='VALID', name=None, **_):
    """tf.nn.conv1d: out[n,t,co] = sum_j sum_ci in[n, t*stride + j, ci] * filters[j, ci, co]
    over the (SAME: zero-padded) input; VALID keeps only fully covered positions."""
    assert stride == 1
    k = filters.shape[0]
    if padding == 'SAME':
        total = k - 1
        value = np.pad(value, [(0, 0), (total // 2, total - total // 2), (0, 0)])
    else:
        assert padding == 'VALID'
    t_out = value.shape[1] - k + 1
    out = np.zeros((value.shape[0], t_out, filters.shape[2]), dtype=np.result_type(value, filters))
    for j in range(k):
        out = out + np.einsum('ntc,co->nto', value[:, j:j + t_out, :], filters[j])
    return _t(out)


def _relu(x, **_):
    return _t(np.maximum(x, 0))


def _conv2d_transpose(value, filter, output_shape, strides, padding='SAME', **_):   # noqa: A002 (mirrors tf)
    """tf.nn.conv2d_transpose (NHWC; filter [height, width, out_channels, in_channels]) = the gradient of conv2d
    with respect to its input: every input pixel scatters filter * pixel into the output at its strided
    position (no kernel flip). Restated for what the reference calls (models.py:116-117): height 1, SAME
    padding; with SAME, conv2d pads (k - s) in total when the output length is a multiple of the stride,
    split floor/ceil, and the transpose drops those positions again."""
    n, h, w_in, cin = value.shape
    kh, kw, cout, cin2 = filter.shape
    assert h == 1 and kh == 1 and cin == cin2 and padding == 'SAME' and list(strides[:2]) == [1, 1] and strides[3] == 1
    s = int(strides[2])
    out_w = int(output_shape[2])
    assert tuple(int(v) for v in output_shape) == (n, 1, out_w, cout) and out_w == w_in * s
    pad_total = max(kw - s, 0)
    pad_left = pad_total // 2
    full = np.zeros((n, 1, (w_in - 1) * s + kw, cout), dtype=np.result_type(value, filter))
    for j in range(kw):
        full[:, 0, j:j + (w_in - 1) * s + 1:s, :] += np.einsum('nlc,oc->nlo', value[:, 0], filter[0, j])
    if full.shape[2] < out_w + pad_left:       # kernel narrower than the stride: untouched positions stay zero
        full = np.pad(full, [(0, 0), (0, 0), (0, out_w + pad_left - full.shape[2]), (0, 0)])
    return _t(full[:, :, pad_left:pad_left + out_w, :])


def _moments(x, axes, keep_dims=False, **_):
    """tf.nn.moments: mean and (population) variance over `axes`."""
    axes = tuple(int(a) for a in axes)
    return _t(np.mean(x, axis=axes, keepdims=keep_dims)), _t(np.var(x, axis=axes, keepdims=keep_dims))


nn = types.SimpleNamespace(conv1d=_conv1d, relu=_relu, conv2d_transpose=_conv2d_transpose, moments=_moments)


class _EMA(object):
    def __init__(self, decay=None, **_):
        self.decay = decay

    def apply(self, var_list=None):
        return None

    def average_name(self, var):
        return str(var) + '/ExponentialMovingAverage'


train = types.SimpleNamespace(ExponentialMovingAverage=_EMA)


class _Logistic(object):
    def __init__(self, loc=0., scale=1.):
        assert loc == 0. and scale == 1.

    def sample(self, shape):
        s = _logistic_sample[0]
        if s is None:
            raise RuntimeError('tf_shim: set_logistic_sample() first')
        return _t(np.reshape(s, [int(v) for v in shape]))


from . import contrib  # noqa: E402  (tensorflow.contrib.{distributions,signal})

contrib.distributions.Logistic = _Logistic
