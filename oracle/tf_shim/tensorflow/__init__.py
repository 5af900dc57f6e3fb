"""ORACLE (test infrastructure, never shipped): a tiny EAGER numpy stand-in for the TensorFlow-1.x API subset
that the reference's hot path touches, so that the reference's OWN Python
(/root/reference/lib_yolo/{layers,model,yolov3,darknet}.py, inference_*.py: concat_bbox / nms) can be executed
in this container, where TensorFlow cannot be installed.

What this pins: the reference's wiring - topology, route indices, channel splits, decode formulas, row layout,
concat order, the NMS call - is executed from the reference's source files unchanged.  What it cannot pin: the
primitive-op semantics of TF itself (marked [TF]); each primitive here is written from
TF's documented behaviour, independently of oracle/net.py (numpy im2col + matmul instead of torch conv, a heap
based NMS instead of a sort).  Variables are not random-initialised: weights, BN statistics and dropout masks
are served by a provider object installed with `set_provider()` in creation order.

Usage (tests/golden/gen_golden.py): put oracle/tf_shim and /root/reference on sys.path, `import tensorflow`.
"""
import contextlib
import heapq

import numpy as np

WORK_DTYPE = np.float32          # gen_golden may switch to float64 for the high-precision goldens
float32 = 'float32'
float64 = 'float64'
int32 = np.int32
int64 = np.int64
uint8 = np.uint8
string = str
bool = np.bool_                   # noqa: A001  (tf.bool)

_PROVIDER = None
_SCOPE = []


def set_provider(p):
    global _PROVIDER
    _PROVIDER = p


def set_work_dtype(dt):
    global WORK_DTYPE
    WORK_DTYPE = dt


def _dt(dtype):
    return WORK_DTYPE if dtype in (float32, None) else (np.float64 if dtype == float64 else dtype)


class TensorShape(list):
    def as_list(self):
        return list(self)


def _unwrap(x):
    if isinstance(x, Tensor):
        return x.a
    if isinstance(x, (list, tuple)):
        return type(x)(_unwrap(v) for v in x)
    return x


class Tensor:
    __array_priority__ = 1000

    def __init__(self, a, name=None):
        self.a = np.asarray(a)
        self.name = name or ('/'.join(_SCOPE) + '/T:0')

    @property
    def shape(self):
        return TensorShape(self.a.shape)

    def get_shape(self):
        return self.shape

    def set_shape(self, s):
        pass

    def __getitem__(self, k):
        return Tensor(self.a[_unwrap(k)])

    def __bool__(self):
        return builtins_bool(self.a)

    def __neg__(self):
        return Tensor(-self.a)

    def __repr__(self):
        return 'shimTensor(%r, %s)' % (self.a.shape, self.a.dtype)


import builtins as _b  # noqa: E402
builtins_bool = _b.bool


def _binop(name):
    def f(self, o):
        return Tensor(getattr(self.a, name)(_unwrap(o)))
    return f


for _n in ('__add__', '__radd__', '__sub__', '__rsub__', '__mul__', '__rmul__', '__truediv__', '__rtruediv__',
           '__pow__', '__lt__', '__gt__', '__le__', '__ge__'):
    setattr(Tensor, _n, _binop(_n))


def _t(x):
    return x if isinstance(x, Tensor) else Tensor(np.asarray(x))


# ------------------------------------------------------------------------------------------ scopes
# [TF] variable_scope(None, default_name=d) makes the name unique among its siblings: d, d_1, d_2, ... (one counter per
# parent scope and default name); an explicit name is used as it is.  Variables are recorded under their scope path the
# way tf.layers names them (<scope>/conv2d/kernel, <scope>/batch_normalization/gamma ...), so that the reference's own
# graph-building code tells us the checkpoint variable names (tests/golden/tf_variable_names.json).
_UNIQ = {}
VARIABLES = []          # (name, shape) in creation order


def reset_names():
    _UNIQ.clear()
    del VARIABLES[:]


def _record_variable(leaf, shape):
    VARIABLES.append(('/'.join(_SCOPE + [leaf]), tuple(int(d) for d in shape)))


@contextlib.contextmanager
def variable_scope(name, default_name=None, **kw):
    if name is None:
        key = ('/'.join(_SCOPE), default_name)
        n = _UNIQ.get(key, 0)
        _UNIQ[key] = n + 1
        name = default_name if n == 0 else '%s_%d' % (default_name, n)
    _SCOPE.append(name)
    try:
        yield
    finally:
        _SCOPE.pop()


@contextlib.contextmanager
def name_scope(name, default_name=None, **kw):          # op names only: does not prefix variables
    yield


class _Contrib:
    class layers:
        @staticmethod
        def l2_regularizer(scale):
            return None


contrib = _Contrib


# ------------------------------------------------------------------------------------------ tf.layers
def _conv2d_numpy(x, k, stride, padding):
    kh, kw, cin, cout = k.shape
    if padding == 'SAME':
        assert stride == 1, 'the reference only uses SAME with stride 1 (layers.py:533-539)'
        x = np.pad(x, ((0, 0), (kh // 2, kh // 2), (kw // 2, kw // 2), (0, 0)))
    else:
        assert padding == 'VALID'
    win = np.lib.stride_tricks.sliding_window_view(x, (kh, kw), axis=(1, 2))    # [S,Ho,Wo,C,kh,kw]
    win = win[:, ::stride, ::stride]
    return np.tensordot(win, k, axes=([4, 5, 3], [0, 1, 2]))


class layers:
    @staticmethod
    def conv2d(inputs, filters, kernel_size, strides=1, activation=None, padding='valid', use_bias=True,
               trainable=True, kernel_regularizer=None, bias_regularizer=None):
        assert activation is None
        x = inputs.a
        w = _PROVIDER.next_conv(kernel_size, x.shape[-1], filters, use_bias)
        k = np.asarray(w['kernel'], dtype=WORK_DTYPE)
        assert k.shape == (kernel_size, kernel_size, x.shape[-1], filters), (k.shape, kernel_size, x.shape, filters)
        _record_variable('conv2d/kernel', k.shape)
        if use_bias:
            _record_variable('conv2d/bias', (filters,))
        y = _conv2d_numpy(x, k, strides, padding.upper())
        if use_bias:
            y = y + np.asarray(w['bias'], dtype=WORK_DTYPE)
        return Tensor(y.astype(WORK_DTYPE))

    @staticmethod
    def batch_normalization(inputs, training=False, trainable=True, epsilon=1e-3):
        assert not training
        w = _PROVIDER.next_bn(inputs.a.shape[-1])
        g, b, m, v = (np.asarray(w[k], dtype=WORK_DTYPE) for k in ('gamma', 'beta', 'mean', 'var'))
        for leaf in ('gamma', 'beta', 'moving_mean', 'moving_variance'):       # creation order of tf.layers.BatchNormalization
            _record_variable('batch_normalization/' + leaf, g.shape)
        inv = g / np.sqrt(v + WORK_DTYPE(epsilon))         # nn.batch_normalization: inv = rsqrt(var+eps)*gamma
        return Tensor(inputs.a * inv + (b - m * inv))

    @staticmethod
    def dropout(inputs, rate=0.5, training=False):
        if not training:
            return inputs
        keep = _PROVIDER.next_dropout_mask(inputs.a.shape, rate)
        return Tensor(inputs.a * WORK_DTYPE(1.0 / (1.0 - rate)) * keep.astype(WORK_DTYPE))

    @staticmethod
    def flatten(t):
        return Tensor(t.a.reshape(t.a.shape[0], -1))


class nn:
    @staticmethod
    def leaky_relu(x, alpha=0.2):
        return Tensor(np.maximum(x.a * WORK_DTYPE(alpha), x.a), name='/'.join(_SCOPE) + '/LeakyRelu:0')

    @staticmethod
    def softmax(x):
        e = np.exp(x.a - np.max(x.a, axis=-1, keepdims=True))
        return Tensor(e / np.sum(e, axis=-1, keepdims=True))


# ------------------------------------------------------------------------------------------ array ops
def shape(x):
    return Tensor(np.asarray(x.a.shape, dtype=np.int32))


def identity(x):
    return Tensor(x.a)


def constant(v, dtype=None):
    return Tensor(np.asarray(v))


def concat(values, axis):
    return Tensor(np.concatenate([_t(v).a for v in values], axis=axis))


def stack(values, axis=0):
    return Tensor(np.stack([_t(v).a for v in values], axis=axis))


def split(value, num_or_size_splits, axis=0):
    a = value.a
    if isinstance(num_or_size_splits, int):
        return [Tensor(p) for p in np.split(a, num_or_size_splits, axis=axis)]
    idx = np.cumsum(num_or_size_splits)[:-1]
    assert sum(num_or_size_splits) == a.shape[axis]
    return [Tensor(p) for p in np.split(a, idx, axis=axis)]


def squeeze(x, axis=None):
    if axis is not None:
        axis = tuple(axis) if isinstance(axis, (list, tuple)) else axis
    return Tensor(np.squeeze(x.a, axis=axis))


def expand_dims(x, axis):
    return Tensor(np.expand_dims(_t(x).a, axis))


def reshape(x, shape):                         # noqa: A002
    return Tensor(np.reshape(x.a, _unwrap(shape)))


def pad(x, paddings, mode='CONSTANT'):
    assert mode == 'CONSTANT'
    return Tensor(np.pad(x.a, paddings))


def gather(params, indices, axis=0):
    return Tensor(np.take(params.a, _unwrap(indices), axis=axis))


def range(n, dtype=None):                      # noqa: A001
    return Tensor(np.arange(n, dtype=_dt(dtype)))


def meshgrid(x, y):
    a, b = np.meshgrid(x.a, y.a)               # default 'xy' indexing, as tf.meshgrid
    return Tensor(a), Tensor(b)


def ones(shape, dtype=float32):                # noqa: A002
    return Tensor(np.ones(shape, dtype=_dt(dtype)))


def ones_like(x):
    return Tensor(np.ones_like(x.a))


def zeros_like(x, dtype=None):
    return Tensor(np.zeros_like(x.a))


def sigmoid(x):
    return Tensor(1 / (1 + np.exp(-x.a)))


def exp(x):
    return Tensor(np.exp(x.a))


def log(x):
    with np.errstate(divide='ignore', invalid='ignore'):
        return Tensor(np.log(x.a))


def _reduce(fn):
    def f(x, axis=None):
        with np.errstate(invalid='ignore'):
            return Tensor(fn(x.a, axis=axis))
    return f


reduce_mean = _reduce(np.mean)
reduce_sum = _reduce(np.sum)
reduce_prod = _reduce(np.prod)


class linalg:
    @staticmethod
    def det(x):
        return Tensor(np.linalg.det(x.a).astype(x.a.dtype))

    @staticmethod
    def diag_part(x):
        return Tensor(np.diagonal(x.a, axis1=-2, axis2=-1))


def while_loop(cond, body, loop_vars, shape_invariants=None):
    v = list(loop_vars)
    while builtins_bool(_unwrap(cond(*v))):
        v = list(body(*v))
    return v


# ------------------------------------------------------------------------------------------ tf.image
def _iou(b, i, j):
    f = b.dtype.type
    ymin_i, xmin_i, ymax_i, xmax_i = min(b[i, 0], b[i, 2]), min(b[i, 1], b[i, 3]), max(b[i, 0], b[i, 2]), max(b[i, 1], b[i, 3])
    ymin_j, xmin_j, ymax_j, xmax_j = min(b[j, 0], b[j, 2]), min(b[j, 1], b[j, 3]), max(b[j, 0], b[j, 2]), max(b[j, 1], b[j, 3])
    area_i = (ymax_i - ymin_i) * (xmax_i - xmin_i)
    area_j = (ymax_j - ymin_j) * (xmax_j - xmin_j)
    if area_i <= 0 or area_j <= 0:
        return f(0)
    inter = max(min(ymax_i, ymax_j) - max(ymin_i, ymin_j), f(0)) * max(min(xmax_i, xmax_j) - max(xmin_i, xmin_j), f(0))
    return inter / (area_i + area_j - inter)


class image:
    @staticmethod
    def resize_nearest_neighbor(x, size):
        a = x.a
        h, w = (int(_unwrap(s)) for s in size)
        assert h == 2 * a.shape[1] and w == 2 * a.shape[2]
        return Tensor(np.repeat(np.repeat(a, 2, axis=1), 2, axis=2))

    @staticmethod
    def non_max_suppression(boxes, scores, max_output_size, iou_threshold=0.5):
        """Priority-queue formulation of TF's NonMaxSuppression kernel; ties pop lower index first."""
        b = np.asarray(boxes.a, dtype=np.float32)
        s = np.asarray(scores.a, dtype=np.float32)
        heap = [(-float(s[i]), i) for i in np.arange(len(s))]
        heapq.heapify(heap)
        sel = []
        thr = np.float32(iou_threshold)
        while heap and len(sel) < max_output_size:
            _, i = heapq.heappop(heap)
            ok = True
            for j in reversed(sel):
                if _iou(b, i, j) > thr:
                    ok = False
                    break
            if ok:
                sel.append(int(i))
        return Tensor(np.asarray(sel, dtype=np.int32))


# things that are only touched when a loss / session / dataset is built: present so imports succeed
class _Missing:
    def __init__(self, n):
        self._n = n

    def __getattr__(self, k):
        return _Missing(self._n + '.' + k)

    def __call__(self, *a, **k):
        raise NotImplementedError('tf shim: %s is outside the hot path' % self._n)


def __getattr__(name):
    return _Missing('tf.' + name)
