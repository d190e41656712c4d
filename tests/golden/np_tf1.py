"""Eager NumPy stand-in for the slice of TensorFlow 1.x / Keras 2 that the REFERENCE's inference
graph touches (TEST INFRASTRUCTURE, build container only).

Purpose: TensorFlow/Keras are not installable offline, so the reference's own source files
(model.py, ops.py, vgg_normalised.py, torchfile.py under /root/reference) cannot run as-is.  This
module registers fake ``tensorflow`` / ``keras`` modules whose ops evaluate immediately on numpy
arrays, so that ``WCTModel(mode='test', ...)`` -- the reference's graph construction code, imported
unmodified -- computes the stylised output while it "builds the graph".  What this pins is the
reference's *algorithm statement* (layer order, reflect padding, transposes, eps placement, k cut,
blending, clip between levels, AdaIN formula): the tensor primitives themselves (conv, svd, pad,
pool) are numpy here, not TensorFlow kernels, and that is said wherever a fixture is used.

placeholder_with_default(name=...) returns the value in FEEDS[name]; the reference's two unnamed
flags (model.py:47 swap5, model.py:51 use_adain) are served from UNNAMED in creation order.
DTYPE selects float32 (reference numerics) or float64 (exact-arithmetic statement of the same code).
"""
import contextlib
import sys
import types

import numpy as np

STATE = dict(dtype=np.float32, feeds={}, unnamed=[], weights={})


def _f(x):
    return np.asarray(x, dtype=STATE["dtype"])


class _Dim(object):
    def __init__(self, v):
        self.value = v


class _Arr(np.ndarray):
    def get_shape(self):
        return tuple(_Dim(int(s)) for s in self.shape)


def _wrap(a):
    return np.asarray(a).view(_Arr)


# ----------------------------------------------------------------------------- tensorflow
def _constant(v, dtype=None, **_):
    a = np.asarray(v)
    return a if a.dtype == np.bool_ else _f(a)


def _placeholder_with_default(default, shape=None, name=None):
    if name is not None:
        if name in STATE["feeds"]:
            v = STATE["feeds"][name]
            return np.asarray(v) if np.asarray(v).dtype == np.bool_ else _f(v)
        return default
    if STATE["unnamed"]:
        v = STATE["unnamed"].pop(0)
        if v is not None:
            return np.asarray(v)
    return default


@contextlib.contextmanager
def _scope(*_a, **_k):
    yield


def _cast(x, dtype):
    if dtype is np.float32:
        return _f(x)
    return np.asarray(x).astype(dtype)


def _svd(x):
    u, s, vt = np.linalg.svd(x)        # tf.svd returns (s, u, v)
    return s, u, vt.T


def _matmul(a, b, transpose_a=False, transpose_b=False):
    a = a.T if transpose_a else a
    b = b.T if transpose_b else b
    return a @ b


def _moments(x, axes, keep_dims=False):
    m = np.mean(x, axis=tuple(axes), keepdims=True)
    v = np.mean((x - m) ** 2, axis=tuple(axes), keepdims=True)
    if not keep_dims:
        m, v = np.squeeze(m, tuple(axes)), np.squeeze(v, tuple(axes))
    return m, v


def _batch_normalization(x, mean, variance, offset, scale, variance_epsilon):
    inv = 1.0 / np.sqrt(variance + STATE["dtype"](variance_epsilon))      # tf.nn.batch_normalization: rsqrt(var+eps)*scale
    inv = inv * scale
    return x * inv + (offset - mean * inv)


def _case(pred_fn_pairs, default):
    for p, fn in pred_fn_pairs:
        if bool(p):
            return fn()
    return default()


def _extract_image_patches(x, ksizes, strides, rates, padding):
    assert padding == "VALID" and list(rates) == [1, 1, 1, 1] and x.shape[0] == 1
    p, st = ksizes[1], strides[1]
    h, w = x.shape[1], x.shape[2]
    rows, cols = (h - p) // st + 1, (w - p) // st + 1
    out = np.empty((1, rows, cols, p * p * x.shape[3]), dtype=x.dtype)
    for r in range(rows):
        for q in range(cols):
            out[0, r, q] = x[0, r * st:r * st + p, q * st:q * st + p, :].reshape(-1)     # (row, col, channel) order
    return out


def _l2_normalize(x, dim, epsilon=1e-12):
    return x / np.sqrt(np.maximum(np.sum(x * x, axis=dim, keepdims=True), STATE["dtype"](epsilon)))


def _conv2d(x, filt, strides, padding):
    assert padding == "VALID" and x.shape[0] == 1
    p, q, cin, cout = filt.shape
    st = strides[1]
    ho, wo = (x.shape[1] - p) // st + 1, (x.shape[2] - q) // st + 1
    f2 = filt.reshape(p * q * cin, cout)
    out = np.empty((1, ho, wo, cout), dtype=x.dtype)
    for y in range(ho):
        for xx in range(wo):
            out[0, y, xx] = x[0, y * st:y * st + p, xx * st:xx * st + q, :].reshape(-1) @ f2
    return out


def _conv2d_transpose(value, filt, output_shape, strides, padding):
    assert padding == "VALID" and value.shape[0] == 1
    p, q, cout, cin = filt.shape                       # [height, width, output_channels, in_channels]
    st = strides[1]
    out = np.zeros(tuple(int(v) for v in output_shape), dtype=value.dtype)
    for y in range(value.shape[1]):
        for x in range(value.shape[2]):
            out[0, y * st:y * st + p, x * st:x * st + q, :] += filt @ value[0, y, x]
    return out


def _deconv_output_length(input_length, filter_size, padding, stride):
    assert padding == "valid"
    return input_length * stride + max(filter_size - stride, 0)


def _make_tf():
    tf = types.ModuleType("tensorflow")
    tf.float32, tf.int32 = np.float32, np.int32
    tf.constant = _constant
    tf.placeholder_with_default = _placeholder_with_default
    tf.name_scope = _scope
    tf.device = _scope
    tf.pad = lambda x, paddings, mode="CONSTANT": np.pad(x, paddings, mode=mode.lower())
    tf.squeeze = lambda x, axis=None: np.squeeze(x, axis)
    tf.transpose = lambda x, perm=None: np.transpose(x, perm)
    tf.shape = lambda x: np.array(np.shape(x), dtype=np.int32)
    tf.unstack = lambda x: [v for v in x]
    tf.reshape = lambda x, shape: np.reshape(x, tuple(int(s) for s in shape))
    tf.reduce_mean = lambda x, axis=None, keep_dims=False: np.mean(x, axis=axis, keepdims=keep_dims)
    tf.reduce_sum = lambda x, axis=None, keep_dims=False: np.sum(x, axis=axis, keepdims=keep_dims)
    tf.matmul = _matmul
    tf.cast = _cast
    tf.eye = lambda n: np.eye(int(n), dtype=STATE["dtype"])
    tf.svd = _svd
    tf.greater = lambda a, b: np.greater(a, b)
    tf.diag = lambda d: np.diag(d)
    tf.pow = lambda x, y: np.power(x, STATE["dtype"](y))
    tf.sqrt = np.sqrt
    tf.expand_dims = lambda x, axis: np.expand_dims(x, axis)
    tf.clip_by_value = lambda x, lo, hi: np.clip(x, lo, hi)
    tf.cond = lambda pred, a, b: a() if bool(pred) else b()
    tf.case = _case
    tf.extract_image_patches = _extract_image_patches
    tf.argmax = lambda x, axis=None: np.argmax(x, axis=axis)
    tf.one_hot = lambda idx, depth, on, off, axis: np.where(np.arange(int(depth)) == np.expand_dims(idx, -1), _f(on), _f(off))
    tf.stack = lambda vals: np.array([int(v) for v in vals])
    tf.ones = lambda shape, dtype=None: np.ones(shape, dtype=STATE["dtype"])
    tf.tile = lambda x, m: np.tile(x, [int(v) for v in m])
    tf.divide = lambda a, b: a / b
    tf.nn = types.SimpleNamespace(moments=_moments, batch_normalization=_batch_normalization, l2_normalize=_l2_normalize,
                                  conv2d=_conv2d, conv2d_transpose=_conv2d_transpose)
    tf.losses = types.SimpleNamespace(mean_squared_error=None)
    py = types.ModuleType("tensorflow.python")
    layers = types.ModuleType("tensorflow.python.layers")
    utils = types.ModuleType("tensorflow.python.layers.utils")
    utils.deconv_output_length = _deconv_output_length
    py.layers, layers.utils, tf.python = layers, utils, py
    return {"tensorflow": tf, "tensorflow.python": py, "tensorflow.python.layers": layers,
            "tensorflow.python.layers.utils": utils}


# ----------------------------------------------------------------------------- keras
class _Sym(object):
    def __init__(self, layer, parent, name=None):
        self.layer, self.parent, self.name = layer, parent, name


class _Layer(object):
    def __init__(self, name=None, **_):
        self.name = name
        self.output = None

    def __call__(self, x):
        if isinstance(x, _Sym):
            self.build_static()
            self.output = _Sym(self, x)
            return self.output
        return self.apply(x)

    def build_static(self):
        pass


class _Lambda(_Layer):
    def __init__(self, fn, name=None, **k):
        _Layer.__init__(self, name)
        self.fn = fn

    def apply(self, x):
        return self.fn(x)


class _Activation(_Layer):
    def __init__(self, kind, name=None, **k):
        _Layer.__init__(self, name)
        assert kind == "relu"

    def apply(self, x):
        return np.maximum(x, 0)


class _MaxPooling2D(_Layer):
    def __init__(self, pool_size=(2, 2), padding="valid", name=None, **k):
        _Layer.__init__(self, name)
        assert padding == "same" and tuple(pool_size) == (2, 2)

    def apply(self, x):
        n, h, w, c = x.shape
        xp = np.pad(x, [(0, 0), (0, h % 2), (0, w % 2), (0, 0)], mode="constant", constant_values=-np.inf)
        return xp.reshape(n, (h + 1) // 2, 2, (w + 1) // 2, 2, c).max(axis=(2, 4))


class _UpSampling2D(_Layer):
    def __init__(self, size=(2, 2), name=None, **k):
        _Layer.__init__(self, name)

    def apply(self, x):
        return np.repeat(np.repeat(x, 2, axis=1), 2, axis=2)


class _Conv2D(_Layer):
    def __init__(self, filters, kernel_size, padding="valid", activation=None, name=None, kernel_initializer=None,
                 bias_initializer=None, trainable=True, **k):
        _Layer.__init__(self, name)
        assert padding == "valid" and activation in (None, "relu")
        self.filters, self.ks, self.activation = int(filters), int(kernel_size), activation
        self.ki, self.bi = kernel_initializer, bias_initializer
        self.kernel = self.bias = None

    def build_static(self):            # Keras runs the initialisers when the layer is first called
        self._resolve()

    def _resolve(self):
        if self.kernel is not None:
            return
        if self.ki is not None:
            self.kernel, self.bias = _f(self.ki(None)), _f(self.bi(None))
        else:                          # a trainable decoder variable: what tf.train.Saver.restore would put there (wct.py:45-56)
            self.kernel, self.bias = (_f(a) for a in STATE["weights"][self.name])
        assert self.kernel.shape[:2] == (self.ks, self.ks) and self.kernel.shape[3] == self.filters, (self.name, self.kernel.shape)

    def apply(self, x):
        self._resolve()
        n, h, w, c = x.shape
        ho, wo = h - self.ks + 1, w - self.ks + 1
        out = np.zeros((n, ho, wo, self.filters), dtype=STATE["dtype"])
        for ky in range(self.ks):
            for kx in range(self.ks):
                out += x[:, ky:ky + ho, kx:kx + wo, :] @ self.kernel[ky, kx]
        out += self.bias
        return np.maximum(out, 0) if self.activation == "relu" else out


def _Input(shape=None, name=None, **k):
    return _Sym(None, None, name)


class _Model(object):
    def __init__(self, inputs=None, outputs=None, name=None):
        self.input, self.outputs, self.name = inputs, outputs, name

    def _eval(self, sym, x, memo):
        if sym is self.input:
            return x
        if id(sym) not in memo:
            memo[id(sym)] = sym.layer.apply(self._eval(sym.parent, x, memo))
        return memo[id(sym)]

    def __call__(self, x):
        assert not isinstance(x, _Sym), "np_tf1 is eager: models are only called on arrays"
        memo = {}
        if isinstance(self.outputs, (list, tuple)):
            return [_wrap(self._eval(o, x, memo)) for o in self.outputs]
        return _wrap(self._eval(self.outputs, x, memo))

    def get_layer(self, name):
        outs = self.outputs if isinstance(self.outputs, (list, tuple)) else [self.outputs]
        for o in outs:
            s = o
            while s is not None and s.layer is not None:
                if s.layer.name == name:
                    return s.layer
                s = s.parent
        raise ValueError("No such layer: " + name)

    def summary(self):
        return ""


def _make_keras():
    keras = types.ModuleType("keras")
    backend = types.ModuleType("keras.backend")
    backend.constant = lambda value, shape=None, **k: _f(value) if shape is None else _f(value).reshape(shape)
    layers = types.ModuleType("keras.layers")
    layers.Input, layers.Conv2D, layers.UpSampling2D = _Input, _Conv2D, _UpSampling2D
    layers.Activation, layers.Lambda, layers.MaxPooling2D = _Activation, _Lambda, _MaxPooling2D
    models = types.ModuleType("keras.models")
    models.Model = _Model
    keras.backend, keras.layers, keras.models = backend, layers, models
    return {"keras": keras, "keras.backend": backend, "keras.layers": layers, "keras.models": models}


REF_DIR = "/root/reference"
_REF_MODULES = ("model", "ops", "vgg_normalised", "torchfile")


@contextlib.contextmanager
def reference_modules():
    """Import the reference's model.py / ops.py / vgg_normalised.py UNMODIFIED over the fake tensorflow/keras and
    restore sys.modules / sys.path afterwards."""
    fakes = {}
    fakes.update(_make_tf())
    fakes.update(_make_keras())
    saved = {k: sys.modules.get(k) for k in list(fakes) + list(_REF_MODULES)}
    sys.modules.update(fakes)
    for m in _REF_MODULES:
        sys.modules.pop(m, None)
    sys.path.insert(0, REF_DIR)
    try:
        import model as ref_model
        import ops as ref_ops
        yield types.SimpleNamespace(model=ref_model, ops=ref_ops)
    finally:
        sys.path.remove(REF_DIR)
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def run_reference(ref, content01, style01, vgg_t7, decoder_weights, relu_targets, alpha, adain, dtype, swap5=False, ss_alpha=0.6,
                  ss_patch_size=3, ss_stride=1):
    """Build (= eagerly evaluate) the reference's WCTModel in test mode exactly as wct.py:31-32 does and return
    (decoded_output of model.py:94, [per-level (content_encoded, decoder_input, decoded)])."""
    import io
    STATE.update(dtype=dtype, feeds={"content_imgs": content01, "style_img": style01, "alpha": alpha, "ss_alpha": ss_alpha},
                 unnamed=[np.bool_(bool(swap5)), np.bool_(bool(adain))], weights=decoder_weights)
    with contextlib.redirect_stdout(io.StringIO()):
        m = ref.model.WCTModel(mode="test", relu_targets=list(relu_targets), vgg_path=vgg_t7,
                               ss_patch_size=ss_patch_size, ss_stride=ss_stride)        # build kwargs, as wct.py:33-36 passes them
    levels = [(np.asarray(e.content_encoded), np.asarray(e.decoder_input), np.asarray(e.decoded)) for e in m.encoder_decoders]
    return np.asarray(m.decoded_output), levels
