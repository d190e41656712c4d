"""Minimal Torch7 (.t7, binary, 8-byte longs) reader for the normalised VGG19 encoder.

Scope row 8f-1: the reference loads ``vgg_normalised.t7`` through its vendored
``torchfile.py`` with ``force_8bytes_long=True`` (vgg_normalised.py:16) and walks
``t7.modules`` (vgg_normalised.py:22-50).  This is an independent re-implementation of just
what that walk needs: the object graph (nil / number / string / boolean / table / torch object
with back-references), Float/Double tensors and storages, and generic ``nn.*`` modules as
attribute dictionaries.  Lua functions are not supported (none occur in the VGG file).
"""
from __future__ import annotations

import struct

import numpy as np

_NIL, _NUMBER, _STRING, _TABLE, _TORCH, _BOOLEAN = 0, 1, 2, 3, 4, 5
_TENSOR_DTYPES = {
    b"torch.FloatTensor": np.float32, b"torch.DoubleTensor": np.float64, b"torch.LongTensor": np.int64,
    b"torch.IntTensor": np.int32, b"torch.ByteTensor": np.uint8, b"torch.CudaTensor": np.float32,
}
_STORAGE_DTYPES = {
    b"torch.FloatStorage": np.float32, b"torch.DoubleStorage": np.float64, b"torch.LongStorage": np.int64,
    b"torch.IntStorage": np.int32, b"torch.ByteStorage": np.uint8, b"torch.CudaStorage": np.float32,
}


class T7Error(ValueError):
    pass


class _Parser(object):
    def __init__(self, data):
        self.b = memoryview(data)
        self.p = 0
        self.memo = {}

    def _take(self, fmt):
        n = struct.calcsize(fmt)
        if self.p + n > len(self.b):
            raise T7Error("truncated .t7 file")
        v = struct.unpack_from(fmt, self.b, self.p)[0]
        self.p += n
        return v

    def i32(self):
        return self._take("<i")

    def i64(self):
        return self._take("<q")

    def raw_string(self):
        n = self.i32()
        s = bytes(self.b[self.p:self.p + n])
        self.p += n
        return s

    def obj(self):
        tag = self.i32()
        if tag == _NIL:
            return None
        if tag == _NUMBER:
            x = self._take("<d")
            return int(x) if float(x).is_integer() else x
        if tag == _BOOLEAN:
            return self.i32() == 1
        if tag == _STRING:
            return self.raw_string()
        if tag not in (_TABLE, _TORCH):
            raise T7Error("unsupported .t7 object tag %d (Lua functions are not supported)" % tag)
        ref = self.i32()
        if ref in self.memo:
            return self.memo[ref]
        if tag == _TABLE:
            out = {}
            self.memo[ref] = out
            for _ in range(self.i32()):
                k = self.obj()
                out[k.decode() if isinstance(k, bytes) else k] = self.obj()
            n = len(out)
            if n and all(isinstance(k, int) for k in out) and sorted(out) == list(range(1, n + 1)):
                lst = [out[i] for i in range(1, n + 1)]          # Lua array -> list
                self.memo[ref] = lst
                return lst
            return out
        # torch object: optional "V <n>" version string, then the class name
        first = self.raw_string()
        cls = self.raw_string() if first.startswith(b"V ") else first
        if cls in _TENSOR_DTYPES:
            nd = self.i32()
            size = [self.i64() for _ in range(nd)]
            stride = [self.i64() for _ in range(nd)]
            off = self.i64() - 1
            storage = self.obj()
            if storage is None or nd == 0:
                arr = np.empty((0,), dtype=_TENSOR_DTYPES[cls])
            else:
                item = storage.dtype.itemsize
                arr = np.lib.stride_tricks.as_strided(storage[off:], shape=size, strides=[s * item for s in stride])
            self.memo[ref] = arr
            return arr
        if cls in _STORAGE_DTYPES:
            n = self.i64()
            dt = np.dtype(_STORAGE_DTYPES[cls])
            arr = np.frombuffer(self.b, dtype=dt, count=n, offset=self.p).copy()
            self.p += n * dt.itemsize
            self.memo[ref] = arr
            return arr
        mod = {"_typename": cls.decode()}
        self.memo[ref] = mod
        payload = self.obj()
        if isinstance(payload, dict):
            mod.update(payload)
        else:
            mod["_payload"] = payload
        return mod


def load(path):
    """Parse a binary .t7 file into plain Python objects (dict / list / numpy / scalars)."""
    with open(path, "rb") as f:
        return _Parser(f.read()).obj()


def load_vgg_t7(path, deepest="relu5_1"):
    """Walk ``modules`` like vgg_from_t7 (vgg_normalised.py:22-50) and return the engine's ``vgg``
    weight list: [{"name", "weight" (O,I,kH,kW) float32, "bias" (O,)} ...]; module 0 is the 1x1
    'preprocess' conv (vgg_normalised.py:25-26); the walk stops after ``deepest``."""
    net = load(path)
    modules = net.get("modules") if isinstance(net, dict) else None
    if not isinstance(modules, list):
        raise T7Error("%s: not an nn.Sequential with a modules list" % path)
    out = []
    for idx, m in enumerate(modules):
        tn = m.get("_typename", "")
        name = m.get("name")
        name = name.decode() if isinstance(name, bytes) else name
        if idx == 0:
            name = "preprocess"
        if tn == "nn.SpatialConvolution":
            w = np.ascontiguousarray(np.asarray(m["weight"], dtype=np.float32))
            if w.ndim == 2:                                   # legacy flattened weight
                w = w.reshape(int(m["nOutputPlane"]), int(m["nInputPlane"]), int(m["kH"]), int(m["kW"]))
            out.append(dict(name=name, weight=w, bias=np.ascontiguousarray(np.asarray(m["bias"], dtype=np.float32))))
        elif tn not in ("nn.SpatialReflectionPadding", "nn.ReLU", "nn.SpatialMaxPooling"):
            raise NotImplementedError(tn)                     # same behaviour as vgg_normalised.py:45-46
        if name == deepest:
            break
    return out
