"""Test utility: serialise a VGG-like nn.Sequential as a binary Torch7 file (8-byte longs), so that
the reference's own torchfile.py (force_8bytes_long=True) and wct_tf_b200.t7 can both read it."""
import struct

import numpy as np


class _W(object):
    def __init__(self):
        self.out = bytearray()
        self.next_ref = 1

    def i32(self, v): self.out += struct.pack("<i", v)
    def i64(self, v): self.out += struct.pack("<q", v)
    def string(self, s):
        b = s if isinstance(s, bytes) else s.encode()
        self.i32(len(b)); self.out += b

    def obj(self, v):
        if v is None:
            self.i32(0)
        elif isinstance(v, bool):
            self.i32(5); self.i32(1 if v else 0)
        elif isinstance(v, (int, float)):
            self.i32(1); self.out += struct.pack("<d", float(v))
        elif isinstance(v, (str, bytes)):
            self.i32(2); self.string(v)
        elif isinstance(v, np.ndarray):
            self.tensor(v)
        elif isinstance(v, list):
            self.table({i + 1: x for i, x in enumerate(v)})
        elif isinstance(v, dict) and "_typename" in v:
            self.i32(4); self.i32(self._ref()); self.string("V 1"); self.string(v["_typename"])
            self.table({k: x for k, x in v.items() if k != "_typename"})
        elif isinstance(v, dict):
            self.table(v)
        else:
            raise TypeError(type(v))

    def _ref(self):
        r = self.next_ref
        self.next_ref += 1
        return r

    def table(self, d):
        self.i32(3); self.i32(self._ref()); self.i32(len(d))
        for k, x in d.items():
            self.obj(k); self.obj(x)

    def tensor(self, a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        self.i32(4); self.i32(self._ref()); self.string("V 1"); self.string("torch.FloatTensor")
        self.i32(a.ndim)
        for s in a.shape: self.i64(s)
        for s in a.strides: self.i64(s // 4)
        self.i64(1)                                            # storage offset, 1-based
        self.i32(4); self.i32(self._ref()); self.string("V 1"); self.string("torch.FloatStorage")
        self.i64(a.size); self.out += a.tobytes()


def write_vgg_t7(path, vgg_layers):
    """vgg_layers: [{"name","weight" (O,I,kH,kW),"bias"}...] starting with the 1x1 preprocess conv."""
    pools = {"conv1_2": "pool1", "conv2_2": "pool2", "conv3_4": "pool3", "conv4_4": "pool4"}
    mods = []
    for i, l in enumerate(vgg_layers):
        w = np.asarray(l["weight"], dtype=np.float32)
        conv = {"_typename": "nn.SpatialConvolution", "nInputPlane": w.shape[1], "nOutputPlane": w.shape[0],
                "kH": w.shape[2], "kW": w.shape[3], "dW": 1, "dH": 1, "padW": 0, "padH": 0,
                "weight": w, "bias": np.asarray(l["bias"], dtype=np.float32), "train": False}
        if i == 0:
            mods.append(conv)                                  # module 0 has no name in the real file either
            continue
        conv["name"] = l["name"]
        mods.append({"_typename": "nn.SpatialReflectionPadding", "pad_l": 1, "pad_r": 1, "pad_t": 1, "pad_b": 1})
        mods.append(conv)
        mods.append({"_typename": "nn.ReLU", "name": l["name"].replace("conv", "relu"), "inplace": True})
        if l["name"] in pools:
            mods.append({"_typename": "nn.SpatialMaxPooling", "name": pools[l["name"]], "kW": 2, "kH": 2, "dW": 2, "dH": 2,
                         "ceil_mode": True})
    w = _W()
    w.obj({"_typename": "nn.Sequential", "modules": mods, "train": False})
    with open(path, "wb") as f:
        f.write(bytes(w.out))
