"""Weight containers for the WCT engine, in the REFERENCE's layouts.

The reference loads
  * the shared encoder from ``vgg_normalised.t7`` (vgg_normalised.py:16-38): conv
    weights (O,I,kH,kW) float32 + bias (O,), transposed to (kH,kW,I,O) at
    vgg_normalised.py:33;
  * one TF checkpoint per decoder (wct.py:47-56): Keras Conv2D kernels
    (3,3,Cin,Cout) + bias (Cout,), layers named ``{relu}_{count}`` (model.py:288).

Neither file exists offline, so benchmarks and parity tests use seeded
synthetic weights in exactly those layouts (``make_synthetic_weights``).

A weights dict is
  {"vgg":      [ {"name", "weight" (O,I,kH,kW), "bias" (O,)} ... in module order ],
   "decoders": { relu: [ {"name", "kernel" (3,3,Cin,Cout), "bias" (Cout,)} ... ] } }
"""
from __future__ import annotations

import numpy as np

from .model import VGG_CONVS, decoder_plan, RELU_TARGETS_ALL

# BGR channel means of the caffe VGG the normalised .t7 was converted from
# (vgg_normalised.py:26 comment: "multiply by 255 and subtract BGR mean as bias").
_BGR_MEAN = np.array([103.939, 116.779, 123.68], dtype=np.float64)


def make_synthetic_weights(seed=42, relu_targets=RELU_TARGETS_ALL, act_rms=1.4,
                           bias_gain=0.3, dec_gain=0.22):
    """Seeded synthetic weights in the reference layouts.

    He-normal 3x3 filters made zero-sum per output filter (no common-mode drive,
    so no channel dies) plus a positive bias ``bias_gain*act_rms``: every relu map
    stays O(1) at every depth (the 'normalised VGG' convention) and the feature
    covariances are well conditioned (no eigenvalue near the hard 1e-5 cut of
    ops.py:112 -- see SURVEY 8c; tests/golden/make_golden.py asserts the gap).
    conv0 is the documented 1x1 preprocess conv; each decoder ends in a
    3-filter conv landing around 0.5 +- ``dec_gain``.  Deterministic in ``seed``."""
    rng = np.random.default_rng(seed)
    vgg = []
    # conv0 "preprocess": x*255, RGB->BGR, minus BGR mean (vgg_normalised.py:25-26)
    w0 = np.zeros((3, 3, 1, 1), dtype=np.float32)
    for o in range(3):
        w0[o, 2 - o, 0, 0] = 255.0
    vgg.append(dict(name="preprocess", weight=w0, bias=(-_BGR_MEAN).astype(np.float32)))
    for name, cin, cout in VGG_CONVS:
        std = np.sqrt(2.0 / (9 * cin))
        if name == "conv1_1":
            std *= act_rms / 80.0  # input ~ 255*U(0,1) - mean: rms ~ 80
        w = rng.normal(0.0, std, size=(cout, cin, 3, 3))
        if name != "conv1_1":
            w -= w.mean(axis=(1, 2, 3), keepdims=True)
        b = np.full(cout, bias_gain * act_rms) + rng.normal(0.0, 0.01, size=cout)
        vgg.append(dict(name=name, weight=w.astype(np.float32), bias=b.astype(np.float32)))
    decoders = {}
    for relu in relu_targets:
        layers = []
        for op in decoder_plan(relu):
            if op.kind != "conv":
                continue
            if op.act:
                k = rng.normal(0.0, np.sqrt(2.0 / (9 * op.cin)), size=(3, 3, op.cin, op.cout))
                b = np.full(op.cout, bias_gain * act_rms) + rng.normal(0.0, 0.01, size=op.cout)
            else:  # final 3-channel conv: land around [0,1]
                k = rng.normal(0.0, dec_gain / (act_rms * np.sqrt(9 * op.cin)), size=(3, 3, op.cin, op.cout))
                b = np.full(op.cout, 0.5)
            k -= k.mean(axis=(0, 1, 2), keepdims=True)
            layers.append(dict(name=op.name, kernel=k.astype(np.float32), bias=b.astype(np.float32)))
        decoders[relu] = layers
    return dict(vgg=vgg, decoders=decoders)


def save_weights(path, weights):
    """Write a weights dict as one ``.npz`` bundle (the engine's own offline format)."""
    flat = {}
    for l in weights["vgg"]:
        flat["vgg/%s/weight" % l["name"]] = l["weight"]
        flat["vgg/%s/bias" % l["name"]] = l["bias"]
    for relu, layers in weights["decoders"].items():
        for l in layers:
            flat["dec/%s/%s/kernel" % (relu, l["name"])] = l["kernel"]
            flat["dec/%s/%s/bias" % (relu, l["name"])] = l["bias"]
    np.savez(path, **flat)


def _load_npz(path):
    z = np.load(path)
    vgg_names = ["preprocess"] + [n for n, _, _ in VGG_CONVS]
    vgg = [dict(name=n, weight=z["vgg/%s/weight" % n], bias=z["vgg/%s/bias" % n])
           for n in vgg_names if "vgg/%s/weight" % n in z.files]
    decoders = {}
    for key in z.files:
        if key.startswith("dec/") and key.endswith("/kernel"):
            _, relu, name, _ = key.split("/")
            decoders.setdefault(relu, []).append(dict(name=name, kernel=z[key], bias=z["dec/%s/%s/bias" % (relu, name)]))
    for relu in decoders:
        decoders[relu].sort(key=lambda l: int(l["name"].rsplit("_", 1)[1]))
    return dict(vgg=vgg, decoders=decoders)


def load_weights(vgg_path, checkpoints, relu_targets):
    """Mirror of the reference's loading protocol (wct.py:47-58): ``checkpoints[i]``
    pairs with ``relu_targets[i]``; a target without a checkpoint raises.

    Encoder: the reference's own Torch7 ``vgg_normalised.t7`` (read by ``t7.py``) or an ``.npz``
    bundle.  Decoders: a TF1 Saver checkpoint directory / prefix as the reference uses (wct.py:51-56;
    read without TensorFlow by ``tf_checkpoint.py`` -- tested only against bundles written by
    tests/tf_bundle_writer.py, never against a file written by TensorFlow itself), or an ``.npz``
    bundle written by ``save_weights`` (each checkpoint holds at least its own decoder)."""
    if vgg_path is None or checkpoints is None:
        raise ValueError("vgg_path and checkpoints are required when no weights dict is given")
    if str(vgg_path).endswith(".t7"):
        from .t7 import load_vgg_t7                      # the reference's own encoder format (vgg_normalised.py:16)
        vgg = load_vgg_t7(vgg_path, deepest=sorted(relu_targets)[-1])
    elif str(vgg_path).endswith(".npz"):
        vgg = _load_npz(vgg_path)["vgg"]
    else:
        raise NotImplementedError("encoder weights must be a Torch7 .t7 file or an .npz bundle")
    decoders = {}
    for relu, ck in zip(relu_targets, checkpoints):     # wct.py:47 zip pairing
        if str(ck).endswith(".npz"):
            d = _load_npz(ck)["decoders"]
            if relu not in d:
                raise Exception('No checkpoint found for target {} in dir {}'.format(relu, ck))  # wct.py:58
            decoders[relu] = d[relu]
        else:
            from .tf_checkpoint import load_decoder_checkpoint   # raises the same Exception when nothing is found
            decoders[relu] = load_decoder_checkpoint(str(ck), relu)
    return dict(vgg=vgg, decoders=decoders)
