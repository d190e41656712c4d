"""CPU restatement of the reference encoder / decoders / level wiring
(TEST INFRASTRUCTURE).

Pinned by tests/golden/pipeline_*.npz: outputs of the reference's OWN model.py / ops.py /
vgg_normalised.py / torchfile.py, imported unmodified and evaluated over a NumPy stand-in for the
TensorFlow/Keras calls they make (tests/golden/np_tf1.py; TensorFlow itself is not installable
offline, so the tensor primitives conv/svd/pad/pool are numpy, the algorithm statement is the
reference's).  tests/test_oracle.py: this module == that code to 1e-9 in float64.

Follows:
  vgg_from_t7      /root/reference/vgg_normalised.py:10-55  (+ ops.py:12-15 pad_reflect)
  build_decoder    /root/reference/model.py:245-304          (+ ops.py:17-19 Conv2DReflect)
  WCTModel wiring  /root/reference/model.py:60-94, 144-158
  WCT.predict      /root/reference/wct.py:60-68, 70-106

Convolutions run through torch-CPU ``F.conv2d`` (cross-correlation, no flip --
same as Keras Conv2D) in the dtype of the input (float32 mimics the reference,
float64 gives a "truth" run).  Tensors at the interface are NHWC numpy arrays
like the reference graph.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import ref_ops

# The module sequence of vgg_normalised.t7 as walked by vgg_normalised.py:22-50.
# (type, name).  'pad' = nn.SpatialReflectionPadding (always 1 px), conv0 is the
# 1x1 "preprocess" conv (vgg_normalised.py:25-26).
VGG_MODULES = [("conv", "preprocess")]
for _blk, _n in [(1, 2), (2, 2), (3, 4), (4, 4), (5, 1)]:
    for _i in range(1, _n + 1):
        VGG_MODULES += [("pad", None), ("conv", "conv%d_%d" % (_blk, _i)), ("relu", "relu%d_%d" % (_blk, _i))]
    if _blk < 5:
        VGG_MODULES += [("pool", "pool%d" % _blk)]

RELU_LEVEL = {"relu1_1": 1, "relu2_1": 2, "relu3_1": 3, "relu4_1": 4, "relu5_1": 5}  # model.py:252

# model.py:255-277 -- ('conv', filters) / ('up',)
DECODER_ARCHS = {
    5: [("conv", 512), ("up",), ("conv", 512), ("conv", 512), ("conv", 512)],
    4: [("conv", 256), ("up",), ("conv", 256), ("conv", 256), ("conv", 256)],
    3: [("conv", 128), ("up",), ("conv", 128)],
    2: [("conv", 64), ("up",)],
    1: [("conv", 64)],
}


def decoder_layers(relu_target):
    """Layer list of one decoder with the reference's names (model.py:283-298):
    count runs over convs AND upsamples; last layer = 3-filter conv, no activation."""
    num = RELU_LEVEL[relu_target]
    out, count = [], 0
    for d in reversed(range(1, num + 1)):
        for tup in DECODER_ARCHS[d]:
            name = "%s_%d" % (relu_target, count)
            if tup[0] == "conv":
                out.append(("conv", name, tup[1], True))
            else:
                out.append(("up", name, None, None))
            count += 1
    out.append(("conv", "%s_%d" % (relu_target, count), 3, False))
    return out


def _t(x_nhwc, dtype):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x_nhwc, dtype=dtype))).permute(0, 3, 1, 2).contiguous()


def _n(x_nchw):
    return x_nchw.permute(0, 2, 3, 1).contiguous().numpy()


def _conv(x, w_hwio, b, pad):
    """Keras Conv2D 'valid' (cross-correlation) on optionally reflect-padded input.
    w_hwio: (kH,kW,I,O) (vgg_normalised.py:33 / Keras kernel layout)."""
    w = torch.from_numpy(np.ascontiguousarray(np.transpose(w_hwio, (3, 2, 0, 1)))).to(x.dtype)
    bb = torch.from_numpy(np.asarray(b)).to(x.dtype)
    if pad:
        x = F.pad(x, (1, 1, 1, 1), mode="reflect")  # ops.py:12-15
    return F.conv2d(x, w, bb)


def encode(img_nhwc, weights, targets, dtype=np.float32):
    """vgg_normalised.py:22-50: run the shared encoder on an NHWC [0,1] image and
    return {relu_name: NHWC feature} for every name in ``targets``."""
    targets = list(targets)
    deepest = sorted(targets)[-1]  # model.py:60
    vgg = {l["name"]: l for l in weights["vgg"]}
    x = _t(img_nhwc, dtype)
    feats = {}
    pad_next = False
    for typ, name in VGG_MODULES:
        if typ == "pad":
            pad_next = True
        elif typ == "conv":
            l = vgg[name]
            w_hwio = np.transpose(np.asarray(l["weight"]), (2, 3, 1, 0))  # vgg_normalised.py:33
            x = _conv(x, w_hwio, l["bias"], pad_next)
            pad_next = False
        elif typ == "relu":
            x = torch.relu(x)
        elif typ == "pool":
            x = F.max_pool2d(x, 2, 2, ceil_mode=True)  # MaxPooling2D(padding='same'), vgg_normalised.py:42
        if name in targets:
            feats[name] = _n(x)
        if name == deepest:  # vgg_normalised.py:48-50
            break
    return feats


def decode(feat_nhwc, weights, relu_target, dtype=np.float32):
    """model.py:279-300: run the ``relu_target`` decoder on an NHWC feature."""
    layers = {l["name"]: l for l in weights["decoders"][relu_target]}
    x = _t(feat_nhwc, dtype)
    for typ, name, _filters, relu in decoder_layers(relu_target):
        if typ == "conv":
            l = layers[name]
            x = _conv(x, np.asarray(l["kernel"]), l["bias"], True)
            if relu:
                x = torch.relu(x)
        else:
            x = x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)  # UpSampling2D, model.py:293
    return _n(x)


def preprocess(image):
    """wct.py:60-64"""
    image = np.asarray(image)
    if image.ndim == 3:
        image = image[None]
    return image / 255.0


def postprocess(image):
    """wct.py:66-68 (np.uint8 truncates)"""
    return np.uint8(np.clip(image, 0, 1) * 255)


def pipeline(content_u8, style_u8, weights, relu_targets, alpha=1.0, adain=False,
             semantics="tf", dtype=np.float32, return_info=False, swap5=False, ss_alpha=0.6, ss_patch_size=3, ss_stride=1):
    """model.py:60-94 + wct.py:70-106 on the CPU.

    semantics: 'tf' -> ops.wct_tf (what the reference graph executes, model.py:154,158)
               'np' -> ops.wct_np (the named oracle)
    Returns the float [0,1]-ish image 1xHxWx3 BEFORE uint8 quantisation
    (= ``decoded_output`` of model.py:94, unclipped)."""
    relu_targets = list(relu_targets)
    content = preprocess(content_u8).astype(dtype)
    style = preprocess(style_u8).astype(dtype)
    style_feats = encode(style, weights, relu_targets, dtype)  # model.py:70-72 (one pass, all targets)
    info = []
    x = content
    for i, relu in enumerate(relu_targets):  # model.py:78
        if i > 0:
            x = np.clip(x, 0, 1)  # model.py:86
        cf = encode(x, weights, [relu], dtype)[relu]  # model.py:135-139
        sf = style_feats[relu]
        if swap5 and relu == "relu5_1":  # model.py:148-152: tf.case gives style-swap precedence over AdaIN at relu5_1
            f, inf = ref_ops.wct_style_swap(cf, sf, ss_alpha, ss_patch_size, ss_stride, return_info=True)
            inf["relu"] = relu
            info.append(inf)
        elif adain:  # model.py:153,157
            f = ref_ops.adain(cf, sf, alpha)
            info.append(dict(relu=relu))
        else:
            fn = ref_ops.wct_tf if semantics == "tf" else ref_ops.wct_np
            f, inf = fn(cf, sf, alpha, return_info=True)
            inf["relu"] = relu
            info.append(inf)
        x = decode(np.asarray(f, dtype=dtype), weights, relu, dtype)  # model.py:173
    if return_info:
        return x, info
    return x
