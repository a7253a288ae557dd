"""Minimal numpy stand-in for the TensorFlow-2 eager API the reference calls.  TEST INFRASTRUCTURE ONLY.

Purpose: TensorFlow/Keras cannot be installed in the build container, so the reference's OWN Python modules
(/root/reference/models/*.py, utils/pipeline.py, layers/_misc.py) are imported on top of this package by
tests/golden/make_golden.py and executed; their outputs become the golden vectors the oracle is pinned to.
Only the primitives those modules call are provided, each with TensorFlow's documented semantics:

  * eager tensors are immutable: `x += y` rebinds (EagerTensor.__iadd__ returns a NEW array) — this matters for
    models/transformer.py:185-190 where `out = baseline; out += mha` must not modify `baseline`;
  * float32 arithmetic throughout; softmax = exp(x - max) / sum; LayerNormalization: biased variance, eps inside
    the sqrt; Conv2D 'same' padding = TF SAME (extra padding on the bottom/right); MaxPooling2D default 2x2 valid;
  * tf.math.top_k: descending, equal values ordered by lower index first;
  * a small functional API (Input / Layer.__call__ on symbolic tensors / Model / get_layer / layers[i].output) with
    Keras' automatic layer names (conv2d, conv2d_1, ...), enough for models/retinanet.py:266-304;
  * tf.keras.applications.mobilenet_v2.MobileNetV2 written with those layers from the upstream definition
    (keras-applications 1.0.8) because the reference hard-wires it (retinanet.py:274).

No file of the reference is copied; nothing here is imported by the product or by the GPU tests.
"""
from __future__ import annotations

import collections
import math as _math
import re as _re
import types as _types

import numpy as np

float32 = np.float32
float64 = np.float64
int32 = np.int32
int64 = np.int64
bool = np.bool_   # noqa: A001
newaxis = None
__version__ = "2.0-numpy-shim"


# ------------------------------------------------------------------------------------------------ eager tensor
class EagerTensor(np.ndarray):
    """ndarray with TensorFlow's value semantics for augmented assignment."""

    def __new__(cls, a):
        return np.asarray(a).view(cls)

    def numpy(self):
        return np.asarray(self)

    def __iadd__(self, o):
        return _t(np.add(np.asarray(self), np.asarray(o)))

    def __isub__(self, o):
        return _t(np.subtract(np.asarray(self), np.asarray(o)))

    def __imul__(self, o):
        return _t(np.multiply(np.asarray(self), np.asarray(o)))

    def __itruediv__(self, o):
        return _t(np.true_divide(np.asarray(self), np.asarray(o)))

    def __hash__(self):
        return id(self)


def _t(a):
    return EagerTensor(a)


def _f32(a):
    a = np.asarray(a)
    return a.astype(np.float32) if a.dtype == np.float64 else a


def convert_to_tensor(x, dtype=None):
    a = np.asarray(x)
    if dtype is not None:
        a = a.astype(dtype)
    return _t(a)


def constant(x, dtype=None):
    a = np.asarray(x)
    if dtype is None and a.dtype == np.int64:
        a = a.astype(np.int32)
    if dtype is None and a.dtype == np.float64:
        a = a.astype(np.float32)
    return convert_to_tensor(a, dtype)


def cast(x, dtype):
    return _t(np.asarray(x).astype(dtype))


def ones(shape, dtype=float32):
    return _t(np.ones(tuple(int(s) for s in shape), dtype))


def zeros(shape, dtype=float32):
    return _t(np.zeros(tuple(int(s) for s in shape), dtype))


def range(*a):   # noqa: A001
    return _t(np.arange(*[int(v) for v in a]).astype(np.int32))


def shape(x):
    return _t(np.array(np.asarray(x).shape, dtype=np.int32))


def reshape(x, shp):
    return _t(np.reshape(np.asarray(x), tuple(int(s) for s in np.asarray(shp).reshape(-1))))


def transpose(x, perm=None):
    return _t(np.transpose(np.asarray(x), perm))


def expand_dims(x, axis):
    a = np.asarray(x)
    if a.dtype == np.int64:
        a = a.astype(np.int32)
    return _t(np.expand_dims(a, axis))


def squeeze(x, axis=None):
    return _t(np.squeeze(np.asarray(x), axis))


def tile(x, multiples):
    return _t(np.tile(np.asarray(x), tuple(int(m) for m in np.asarray(multiples))))


def stack(xs, axis=0):
    return _t(np.stack([np.asarray(x) for x in xs], axis))


def concat(xs, axis):
    return _t(np.concatenate([np.asarray(x) for x in xs], axis))


def gather_nd(params, indices):
    p, idx = np.asarray(params), np.asarray(indices)
    k = idx.shape[-1]
    flat = idx.reshape(-1, k)
    out = np.stack([p[tuple(r)] for r in flat], 0)
    return _t(out.reshape(idx.shape[:-1] + p.shape[k:]))


def argmax(x, axis=None, output_type=int64):
    return _t(np.argmax(np.asarray(x), axis).astype(output_type))


def maximum(a, b):
    return _t(np.maximum(np.asarray(a), np.asarray(b)))


def minimum(a, b):
    return _t(np.minimum(np.asarray(a), np.asarray(b)))


def matmul(a, b, transpose_a=False, transpose_b=False):
    a, b = np.asarray(a), np.asarray(b)
    if transpose_a:
        a = np.swapaxes(a, -1, -2)
    if transpose_b:
        b = np.swapaxes(b, -1, -2)
    return _t(np.matmul(a, b))


def reduce_mean(x, axis=None, keepdims=False):
    return _t(np.mean(np.asarray(x), axis=axis, keepdims=keepdims))


def function(fn=None, **kw):
    return fn if fn is not None else (lambda f: f)


class GradientTape:   # training only; never executed by the golden generator
    def __enter__(self):
        raise NotImplementedError("training is out of scope for the shim")

    def __exit__(self, *a):
        return False


_TopK = collections.namedtuple("TopKV2", ["values", "indices"])

math = _types.SimpleNamespace(
    equal=lambda a, b: _t(np.equal(np.asarray(a), np.asarray(b))),
    less_equal=lambda a, b: _t(np.less_equal(np.asarray(a), np.asarray(b))),
    logical_not=lambda a: _t(np.logical_not(np.asarray(a))),
    sqrt=lambda a: _t(np.sqrt(np.asarray(a))),
    rsqrt=lambda a: _t(1.0 / np.sqrt(np.asarray(a))),
    square=lambda a: _t(np.square(np.asarray(a))),
    minimum=minimum, maximum=maximum,
    reduce_sum=lambda x, axis=None: _t(np.sum(np.asarray(x), axis=axis)),
    reduce_mean=reduce_mean,
    reduce_max=lambda x, axis=None: _t(np.max(np.asarray(x), axis=axis)),
    reduce_min=lambda x, axis=None: _t(np.min(np.asarray(x), axis=axis)),
)


def _top_k(x, k=1, sorted=True):   # noqa: A002
    a = np.asarray(x)
    assert a.ndim == 1, "shim top_k: 1-D input only (the reference flattens, pipeline.py:123-128)"
    order = np.argsort(-a, kind="stable")[: int(k)]          # descending; ties -> lower index first
    return _TopK(_t(a[order]), _t(order.astype(np.int32)))


math.top_k = _top_k

linalg = _types.SimpleNamespace(
    band_part=lambda x, lo, hi: _t(np.tril(np.asarray(x)) if (lo, hi) == (-1, 0) else _bad("band_part")))


def _bad(what):
    raise NotImplementedError("tensorflow shim: " + what)


def _softmax(x, axis=-1):
    a = np.asarray(x).astype(np.float32)
    e = np.exp(a - a.max(axis=axis, keepdims=True))
    return _t((e / e.sum(axis=axis, keepdims=True)).astype(np.float32))


def _leaky_relu(x, alpha=0.2):
    a = np.asarray(x)
    return _t(np.where(a >= 0, a, a * np.float32(alpha)).astype(a.dtype))


nn = _types.SimpleNamespace(softmax=_softmax, leaky_relu=_leaky_relu,
                            relu=lambda x: _t(np.maximum(np.asarray(x), 0)),
                            relu6=lambda x: _t(np.clip(np.asarray(x), 0, 6)))


# ------------------------------------------------------------------------------------------------ image
class _ResizeMethod:
    BILINEAR, NEAREST_NEIGHBOR, BICUBIC, AREA = "bilinear", "nearest", "bicubic", "area"


def _resize(images, size, method="bilinear", preserve_aspect_ratio=False, antialias=False, name=None):
    """tf.image.resize (TF2, half-pixel centres).  nearest: src = floor((dst + 0.5) * scale)."""
    a = np.asarray(images)
    squeeze_b = a.ndim == 3
    if squeeze_b:
        a = a[None]
    ho, wo = int(size[0]), int(size[1])
    h, w = a.shape[1], a.shape[2]
    if method == "nearest":
        yi = np.minimum(np.floor((np.arange(ho) + 0.5) * (h / ho)).astype(np.int64), h - 1)
        xi = np.minimum(np.floor((np.arange(wo) + 0.5) * (w / wo)).astype(np.int64), w - 1)
        out = a[:, yi][:, :, xi]
    elif method == "bilinear":
        def axis_w(n_in, n_out):
            s = (np.arange(n_out) + 0.5) * (n_in / n_out) - 0.5
            lo = np.floor(s)
            f = (s - lo).astype(np.float32)
            i0 = np.clip(lo, 0, n_in - 1).astype(np.int64)
            i1 = np.clip(lo + 1, 0, n_in - 1).astype(np.int64)
            return i0, i1, f
        y0, y1, fy = axis_w(h, ho)
        x0, x1, fx = axis_w(w, wo)
        af = a.astype(np.float32)
        top = af[:, y0][:, :, x0] * (1 - fx)[None, None, :, None] + af[:, y0][:, :, x1] * fx[None, None, :, None]
        bot = af[:, y1][:, :, x0] * (1 - fx)[None, None, :, None] + af[:, y1][:, :, x1] * fx[None, None, :, None]
        out = top * (1 - fy)[None, :, None, None] + bot * fy[None, :, None, None]
    else:
        _bad("resize method " + str(method))
    return _t(out[0] if squeeze_b else out)


image = _types.SimpleNamespace(ResizeMethod=_ResizeMethod, resize=_resize,
                               decode_jpeg=lambda *a, **k: _bad("decode_jpeg"))
io = _types.SimpleNamespace(read_file=lambda *a, **k: _bad("read_file"))
test = _types.SimpleNamespace(gpu_device_name=lambda: "")
data = _types.SimpleNamespace(experimental=_types.SimpleNamespace(AUTOTUNE=-1), Dataset=object)


# ------------------------------------------------------------------------------------------------ keras
_UIDS = collections.defaultdict(int)


def _snake(name):
    s = _re.sub("(.)([A-Z][a-z0-9]+)", r"\1_\2", name)
    return _re.sub("([a-z])([A-Z])", r"\1_\2", s).lower()


def _auto_name(cls_name):
    base = _snake(cls_name)
    n = _UIDS[base]
    _UIDS[base] += 1
    return base if n == 0 else "%s_%d" % (base, n)


def reset_uids():
    _UIDS.clear()


class KTensor:
    """Symbolic tensor of the functional API: output `index` of `layer` applied to `inputs`."""

    def __init__(self, layer, inputs, index=0):
        self.layer, self.inputs, self.index = layer, inputs, index


def _flatten(x):
    if isinstance(x, (list, tuple)):
        out = []
        for e in x:
            out.extend(_flatten(e))
        return out
    return [x]


def _is_symbolic(x):
    return any(isinstance(e, KTensor) for e in _flatten(x))


class Layer:
    def __init__(self, name=None, trainable=True, dtype=None, **kw):
        self.name = name if name is not None else _auto_name(type(self).__name__)
        self.built = False
        self._vars = collections.OrderedDict()
        self.output = None
        self.input = None

    # variables -------------------------------------------------------------------------------
    def add_var(self, name, shape):
        self._vars[name] = np.zeros(shape, np.float32)

    def set_var(self, name, value):
        assert name in self._vars, "%s has no variable %s" % (self.name, name)
        v = np.asarray(value, np.float32)
        assert v.shape == self._vars[name].shape, (self.name, name, v.shape, self._vars[name].shape)
        self._vars[name] = v

    def assign(self, name, value):
        """Set (creating if necessary) a variable; a layer whose variables were assigned counts as built."""
        self._vars[name] = np.ascontiguousarray(value, np.float32)
        self.built = True

    def build(self, input_shape):
        pass

    def call(self, *a, **k):
        raise NotImplementedError

    def __call__(self, inputs, *args, **kwargs):
        if _is_symbolic(inputs) or any(isinstance(a, KTensor) for a in args):
            sym_in = list(_flatten(inputs)) + [a for a in args if isinstance(a, KTensor)]
            n_in = len(_flatten(inputs))
            extra = [a for a in args if not isinstance(a, KTensor)]
            layer = self

            class _Node:
                pass
            node = _Node()
            node.layer, node.sym_in, node.n_in, node.struct = layer, sym_in, n_in, inputs
            node.extra, node.kwargs, node.nargs_sym = extra, kwargs, len(sym_in) - n_in
            out = KTensor(node, sym_in, 0)
            self.output, self.input = out, inputs
            return out
        return self._eager(inputs, *args, **kwargs)

    def _eager(self, inputs, *args, **kwargs):
        if not self.built:
            first = _flatten(inputs)[0]
            self.build(np.asarray(first).shape)
            self.built = True
        return self.call(inputs, *args, **kwargs)


class InputLayer(Layer):
    def __init__(self, input_shape=None, sparse=False, name=None, **kw):
        super().__init__(name=name if name is not None else _auto_name("input"))

    def call(self, x):
        return x


def Input(shape=None, name=None, **kw):   # noqa: N802
    layer = InputLayer(name=name)
    t = KTensor(None, [], 0)
    t.input_layer = layer
    layer.output = t
    return t


class Dense(Layer):
    def __init__(self, units, activation=None, kernel_initializer=None, **kw):
        super().__init__(**kw)
        self.units, self.activation = units, _activation(activation)

    def build(self, shp):
        self.add_var("kernel", (shp[-1], self.units))
        self.add_var("bias", (self.units,))

    def ensure_built(self, cin):
        if not self.built:
            self.build((None, cin))
            self.built = True

    def call(self, x):
        y = np.matmul(np.asarray(x, np.float32), self._vars["kernel"]) + self._vars["bias"]
        return self.activation(_t(y.astype(np.float32)))


class LayerNormalization(Layer):
    def __init__(self, epsilon=1e-3, **kw):
        super().__init__(**kw)
        self.eps = epsilon

    def build(self, shp):
        self.add_var("gamma", (shp[-1],))
        self.add_var("beta", (shp[-1],))
        self._vars["gamma"] = np.ones((shp[-1],), np.float32)

    def call(self, x):
        a = np.asarray(x, np.float32)
        mean = a.mean(-1, keepdims=True)
        var = ((a - mean) ** 2).mean(-1, keepdims=True)
        y = (a - mean) / np.sqrt(var + np.float32(self.eps)) * self._vars["gamma"] + self._vars["beta"]
        return _t(y.astype(np.float32))


class Dropout(Layer):
    def __init__(self, rate, **kw):
        super().__init__(**kw)

    def call(self, x, training=None):
        assert not training, "shim Dropout: inference only"
        return x


class Embedding(Layer):
    def __init__(self, input_dim, output_dim, **kw):
        super().__init__(**kw)
        self.input_dim, self.output_dim = input_dim, output_dim

    def build(self, shp):
        self.add_var("embeddings", (self.input_dim, self.output_dim))

    def call(self, x):
        return _t(self._vars["embeddings"][np.asarray(x).astype(np.int64)])


def _same_pad(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def _conv2d(x, kernel, strides, padding, groups_depthwise=False):
    """NHWC conv by im2col + matmul in float32.  kernel HWIO (or HWC1 for depthwise)."""
    x = np.asarray(x, np.float32)
    kh, kw = kernel.shape[0], kernel.shape[1]
    s = strides
    if padding == "same":
        pt, pb = _same_pad(x.shape[1], kh, s)
        pl, pr = _same_pad(x.shape[2], kw, s)
        x = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)))
    n, h, w, c = x.shape
    ho, wo = (h - kh) // s + 1, (w - kw) // s + 1
    cols = np.empty((n, ho, wo, kh, kw, c), np.float32)
    for i in np.arange(kh):
        for j in np.arange(kw):
            cols[:, :, :, i, j, :] = x[:, i:i + s * ho:s, j:j + s * wo:s, :]
    if groups_depthwise:
        return np.einsum("nhwijc,ijc->nhwc", cols, kernel[:, :, :, 0]).astype(np.float32)
    return np.matmul(cols.reshape(n * ho * wo, kh * kw * c), kernel.reshape(kh * kw * c, -1)).reshape(n, ho, wo, -1)


def _activation(a):
    if a is None or a == "linear":
        return lambda x: x
    if a == "relu":
        return nn.relu
    if callable(a):
        return a
    _bad("activation " + str(a))


class Conv2D(Layer):
    def __init__(self, filters, kernel_size, strides=1, padding="valid", activation=None, use_bias=True,
                 kernel_initializer=None, bias_initializer=None, **kw):
        super().__init__(**kw)
        ks = kernel_size if isinstance(kernel_size, (tuple, list)) else (kernel_size, kernel_size)
        st = strides[0] if isinstance(strides, (tuple, list)) else strides
        self.filters, self.ks, self.strides, self.padding = filters, tuple(ks), st, padding
        self.use_bias, self.activation = use_bias, _activation(activation)

    def build(self, shp):
        self.add_var("kernel", self.ks + (shp[-1], self.filters))
        if self.use_bias:
            self.add_var("bias", (self.filters,))

    def call(self, x):
        y = _conv2d(x, self._vars["kernel"], self.strides, self.padding)
        if self.use_bias:
            y = y + self._vars["bias"]
        return self.activation(_t(y.astype(np.float32)))


class DepthwiseConv2D(Layer):
    def __init__(self, kernel_size, strides=1, padding="valid", use_bias=True, activation=None, **kw):
        super().__init__(**kw)
        self.ks = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
        self.strides = strides[0] if isinstance(strides, (tuple, list)) else strides
        self.padding, self.use_bias = padding, use_bias

    def build(self, shp):
        self.add_var("depthwise_kernel", self.ks + (shp[-1], 1))

    def call(self, x):
        return _t(_conv2d(x, self._vars["depthwise_kernel"], self.strides, self.padding, groups_depthwise=True))


class BatchNormalization(Layer):
    def __init__(self, axis=-1, epsilon=1e-3, momentum=0.99, **kw):
        super().__init__(**kw)
        self.eps = epsilon

    def build(self, shp):
        c = shp[-1]
        for n in ("gamma", "beta", "moving_mean", "moving_variance"):
            self.add_var(n, (c,))

    def call(self, x, training=None):
        v = self._vars
        y = (np.asarray(x, np.float32) - v["moving_mean"]) / np.sqrt(v["moving_variance"] + np.float32(self.eps))
        return _t((y * v["gamma"] + v["beta"]).astype(np.float32))


class ReLU(Layer):
    def __init__(self, max_value=None, **kw):
        super().__init__(**kw)
        self.max_value = max_value

    def call(self, x):
        a = np.maximum(np.asarray(x), 0)
        return _t(a if self.max_value is None else np.minimum(a, np.float32(self.max_value)))


class Activation(Layer):
    def __init__(self, activation, **kw):
        super().__init__(**kw)
        self.fn = _activation(activation)

    def call(self, x):
        return self.fn(x)


class ZeroPadding2D(Layer):
    def __init__(self, padding=(1, 1), **kw):
        super().__init__(**kw)
        if isinstance(padding, int):
            padding = ((padding, padding), (padding, padding))
        elif isinstance(padding[0], int):
            padding = ((padding[0], padding[0]), (padding[1], padding[1]))
        self.padding = padding

    def call(self, x):
        (t, b), (l, r) = self.padding
        return _t(np.pad(np.asarray(x), ((0, 0), (t, b), (l, r), (0, 0))))


class Add(Layer):
    def call(self, xs):
        out = np.asarray(xs[0])
        for x in xs[1:]:
            out = out + np.asarray(x)
        return _t(out)


class Concatenate(Layer):
    def __init__(self, axis=-1, **kw):
        super().__init__(**kw)
        self.axis = axis

    def call(self, xs):
        return _t(np.concatenate([np.asarray(x) for x in xs], self.axis))


class MaxPooling2D(Layer):
    def __init__(self, pool_size=(2, 2), strides=None, padding="valid", **kw):
        super().__init__(**kw)
        self.pool = (pool_size, pool_size) if isinstance(pool_size, int) else tuple(pool_size)
        self.strides = self.pool if strides is None else ((strides, strides) if isinstance(strides, int) else tuple(strides))
        self.padding = padding

    def call(self, x):
        a = np.asarray(x)
        (ph, pw), (sh, sw) = self.pool, self.strides
        if self.padding == "same":
            pt, pb = _same_pad(a.shape[1], ph, sh)
            pl, pr = _same_pad(a.shape[2], pw, sw)
            a = np.pad(a, ((0, 0), (pt, pb), (pl, pr), (0, 0)), constant_values=-np.inf)
        n, h, w, c = a.shape
        ho, wo = (h - ph) // sh + 1, (w - pw) // sw + 1
        out = np.full((n, ho, wo, c), -np.inf, a.dtype)
        for i in np.arange(ph):
            for j in np.arange(pw):
                out = np.maximum(out, a[:, i:i + sh * ho:sh, j:j + sw * wo:sw, :])
        return _t(out)


class Model(Layer):
    """Functional model (inputs/outputs given) or subclassed model (neither given)."""

    def __init__(self, inputs=None, outputs=None, name=None, **kw):
        super().__init__(name=name if name is not None else _auto_name("model"))
        self.functional = inputs is not None
        if not self.functional:
            return
        self.inputs = inputs if isinstance(inputs, (list, tuple)) else [inputs]
        self.outputs = outputs if isinstance(outputs, (list, tuple)) else [outputs]
        self._single_out = not isinstance(outputs, (list, tuple))
        self._in_flat = _flatten(self.inputs)
        # layers in Keras order: input layers first, then layers by depth-first creation order of the graph
        seen, order, in_layers = set(), [], []

        def visit(t):
            if id(t) in seen:
                return
            seen.add(id(t))
            if t.layer is None:
                if any(t is i for i in self._in_flat) and t.input_layer not in in_layers:
                    in_layers.append(t.input_layer)
                return
            if any(t is i for i in self._in_flat):
                return
            for s in t.layer.sym_in:
                visit(s)
            if t.layer.layer not in order:
                order.append(t.layer.layer)
        for o in _flatten(self.outputs):
            visit(o)
        for i in self._in_flat:      # inputs that are intermediate tensors of another graph
            if i.layer is None and i.input_layer not in in_layers:
                in_layers.append(i.input_layer)
        self.layers = in_layers + order
        self.built = True

    def get_layer(self, name):
        for l in self.layers:
            if l.name == name:
                return l
        raise ValueError("No such layer: " + name)

    def load_weights(self, path):
        # no HDF5 reader here and no weight file ships with the reference: record the request; the golden generator
        # assigns every variable explicitly afterwards (Layer.assign)
        self.requested_weights = path

    def all_layers(self):
        out = []
        for l in self.layers:
            out.append(l)
            if isinstance(l, Model) and l.functional:
                out.extend(l.all_layers())
        return out

    def call(self, inputs, *a, **k):
        assert self.functional
        vals = _flatten(inputs)
        assert len(vals) == len(self._in_flat), "model %s expects %d inputs" % (self.name, len(self._in_flat))
        memo = {id(t): v for t, v in zip(self._in_flat, vals)}

        def ev(t):
            if id(t) in memo:
                return memo[id(t)]
            node = t.layer
            assert node is not None, "graph input of model %s was not fed" % self.name
            ins = [ev(s) for s in node.sym_in]
            main = ins[: node.n_in]
            arg = _rebuild(node.struct, iter(main))
            extra_sym = ins[node.n_in:]
            out = node.layer._eager(arg, *extra_sym, *node.extra, **node.kwargs)
            memo[id(t)] = out
            return out
        outs = [ev(o) for o in _flatten(self.outputs)]
        return outs[0] if self._single_out else outs

    def __call__(self, inputs, *args, **kwargs):
        if not self.functional:
            return self.call(inputs, *args, **kwargs)
        if _is_symbolic(inputs):
            res = Layer.__call__(self, inputs, *args, **kwargs)
            if self._single_out:
                return res
            node = res.layer
            outs = [KTensor(_Pick(node, i), node.sym_in, 0) for i in np.arange(len(self.outputs))]
            return outs
        return self.call(inputs)


class _Pick:
    """Node selecting output i of a multi-output nested model."""

    def __init__(self, node, i):
        self.sym_in, self.n_in, self.struct = node.sym_in, node.n_in, node.struct
        self.extra, self.kwargs = node.extra, node.kwargs
        inner = node.layer

        class _Sel(Layer):
            def _eager(self_, arg, *a, **k):   # noqa: N805
                return inner._eager(arg, *a, **k)[int(i)]
        self.layer = _Sel(name=inner.name + "_out%d" % int(i))


def _rebuild(struct, it):
    if isinstance(struct, (list, tuple)):
        return [_rebuild(s, it) for s in struct]
    return next(it)


class _Initializer:
    def __init__(self, *a, **k):
        pass


class _LearningRateSchedule:
    def __init__(self, *a, **k):
        pass


class _Anything:
    """Placeholder for training-only objects (optimizers, losses, metrics, checkpoints)."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __bool__(self):
        return False


class _Tokenizer:
    """tf.keras.preprocessing.text.Tokenizer == keras_preprocessing.text.Tokenizer (third-party, un-vendored; the TF 2.0/2.1
    the reference targets bundles keras-preprocessing 1.1.0).  Restated from its published algorithm: fit_on_texts,
    texts_to_sequences, sequences_to_texts (num_words / oov_token semantics), get_config / to_json."""

    def __init__(self, num_words=None, filters='!"#$%&()*+,-./:;<=>?@[\\]^_`{|}~\t\n', lower=True, split=" ", char_level=False,
                 oov_token=None, document_count=0, **kwargs):
        import collections
        self.word_counts = collections.OrderedDict()
        self.word_docs = collections.defaultdict(int)
        self.filters, self.split, self.lower, self.num_words = filters, split, lower, num_words
        self.document_count, self.char_level, self.oov_token = document_count, char_level, oov_token
        self.index_docs = collections.defaultdict(int)
        self.word_index, self.index_word = {}, {}

    def _words(self, text):
        if self.lower:
            text = text.lower()
        text = text.translate(str.maketrans({c: self.split for c in self.filters}))
        return [w for w in text.split(self.split) if w]

    def fit_on_texts(self, texts):
        for text in texts:
            self.document_count += 1
            seq = self._words(text)
            for w in seq:
                self.word_counts[w] = self.word_counts.get(w, 0) + 1
            for w in set(seq):
                self.word_docs[w] += 1
        wcounts = list(self.word_counts.items())
        wcounts.sort(key=lambda x: x[1], reverse=True)
        sorted_voc = [] if self.oov_token is None else [self.oov_token]
        sorted_voc.extend(wc[0] for wc in wcounts)
        import builtins
        self.word_index = dict(zip(sorted_voc, list(builtins.range(1, len(sorted_voc) + 1))))   # index 0 is reserved
        self.index_word = {c: w for w, c in self.word_index.items()}
        for w, c in list(self.word_docs.items()):
            self.index_docs[self.word_index[w]] = c

    def texts_to_sequences(self, texts):
        num_words, oov = self.num_words, self.word_index.get(self.oov_token)
        out = []
        for text in texts:
            vect = []
            for w in self._words(text):
                i = self.word_index.get(w)
                if i is not None:
                    if num_words and i >= num_words:
                        if oov is not None:
                            vect.append(oov)
                    else:
                        vect.append(i)
                elif self.oov_token is not None:
                    vect.append(oov)
            out.append(vect)
        return out

    def sequences_to_texts(self, sequences):
        num_words, oov = self.num_words, self.word_index.get(self.oov_token)
        out = []
        for seq in sequences:
            vect = []
            for num in seq:
                word = self.index_word.get(num)
                if word is not None:
                    if num_words and num >= num_words:
                        if oov is not None:
                            vect.append(self.index_word[oov])
                    else:
                        vect.append(word)
                elif self.oov_token is not None:
                    vect.append(self.index_word[oov])
            out.append(" ".join(vect))
        return out

    def get_config(self):
        import json
        return {"num_words": self.num_words, "filters": self.filters, "lower": self.lower, "split": self.split,
                "char_level": self.char_level, "oov_token": self.oov_token, "document_count": self.document_count,
                "word_counts": json.dumps(self.word_counts), "word_docs": json.dumps(self.word_docs),
                "index_docs": json.dumps(self.index_docs), "index_word": json.dumps(self.index_word),
                "word_index": json.dumps(self.word_index)}

    def to_json(self, **kwargs):
        import json
        return json.dumps({"class_name": self.__class__.__name__, "config": self.get_config()}, **kwargs)


def _mobilenet_v2(input_tensor=None, alpha=1.0, include_top=False, pooling=None, weights=None, **kw):
    """Keras-applications MobileNetV2 (alpha = 1.0), include_top=False; layer names as upstream."""
    assert alpha == 1.0 and not include_top and weights is None
    L = layers

    def bn(name):
        return L.BatchNormalization(epsilon=1e-3, momentum=0.999, name=name)

    x = L.ZeroPadding2D(padding=((0, 1), (0, 1)), name="Conv1_pad")(input_tensor)
    x = L.Conv2D(32, 3, strides=2, padding="valid", use_bias=False, name="Conv1")(x)
    x = bn("bn_Conv1")(x)
    x = L.ReLU(6.0, name="Conv1_relu")(x)
    cin = 32

    def block(x, cin, cout, stride, expansion, bid):
        prefix = "block_%d_" % bid if bid else "expanded_conv_"
        inp = x
        if bid:
            x = L.Conv2D(expansion * cin, 1, padding="same", use_bias=False, name=prefix + "expand")(x)
            x = bn(prefix + "expand_BN")(x)
            x = L.ReLU(6.0, name=prefix + "expand_relu")(x)
        if stride == 2:
            x = L.ZeroPadding2D(padding=((0, 1), (0, 1)), name=prefix + "pad")(x)
        x = L.DepthwiseConv2D(3, strides=stride, use_bias=False, padding="same" if stride == 1 else "valid",
                              name=prefix + "depthwise")(x)
        x = bn(prefix + "depthwise_BN")(x)
        x = L.ReLU(6.0, name=prefix + "depthwise_relu")(x)
        x = L.Conv2D(cout, 1, padding="same", use_bias=False, name=prefix + "project")(x)
        x = bn(prefix + "project_BN")(x)
        if cin == cout and stride == 1:
            return L.Add(name=prefix + "add")([inp, x])
        return x

    x = block(x, 32, 16, 1, 1, 0)
    cin = 16
    cfg = [(24, 2), (24, 1), (32, 2), (32, 1), (32, 1), (64, 2), (64, 1), (64, 1), (64, 1), (96, 1), (96, 1), (96, 1),
           (160, 2), (160, 1), (160, 1), (320, 1)]
    for i, (cout, stride) in enumerate(cfg):
        x = block(x, cin, cout, stride, 6, i + 1)
        cin = cout
    x = L.Conv2D(1280, 1, use_bias=False, name="Conv_1")(x)
    x = bn("Conv_1_bn")(x)
    x = L.ReLU(6.0, name="out_relu")(x)
    return Model(inputs=input_tensor, outputs=x, name="mobilenetv2_1.00_224")


layers = _types.SimpleNamespace(
    Layer=Layer, InputLayer=InputLayer, Input=Input, Dense=Dense, LayerNormalization=LayerNormalization, Dropout=Dropout,
    Embedding=Embedding, Conv2D=Conv2D, DepthwiseConv2D=DepthwiseConv2D, BatchNormalization=BatchNormalization, ReLU=ReLU,
    Activation=Activation, ZeroPadding2D=ZeroPadding2D, Add=Add, Concatenate=Concatenate, MaxPooling2D=MaxPooling2D,
    MaxPool2D=MaxPooling2D)

keras = _types.SimpleNamespace(
    layers=layers, Model=Model, models=_types.SimpleNamespace(Model=Model),
    initializers=_types.SimpleNamespace(he_normal=lambda *a, **k: _Initializer(), RandomNormal=_Initializer,
                                        Initializer=_Initializer),
    backend=_types.SimpleNamespace(shape=shape, image_data_format=lambda: "channels_last"),
    applications=_types.SimpleNamespace(mobilenet_v2=_types.SimpleNamespace(
        MobileNetV2=_mobilenet_v2,
        preprocess_input=lambda x: _t(np.asarray(x, np.float32) / np.float32(127.5) - np.float32(1.0)))),
    optimizers=_types.SimpleNamespace(Adam=_Anything, schedules=_types.SimpleNamespace(LearningRateSchedule=_LearningRateSchedule)),
    losses=_types.SimpleNamespace(SparseCategoricalCrossentropy=_Anything, MeanSquaredError=_Anything),
    metrics=_types.SimpleNamespace(Mean=_Anything),
    preprocessing=_types.SimpleNamespace(text=_types.SimpleNamespace(Tokenizer=_Tokenizer),
                                         sequence=_types.SimpleNamespace(pad_sequences=lambda *a, **k: _bad("pad_sequences"))),
)

train = _types.SimpleNamespace(Checkpoint=_Anything, CheckpointManager=_Anything)
