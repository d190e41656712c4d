"""CPU experiment (TEST INFRASTRUCTURE, run by hand): where does the free-running 5-level noise come from?

VERDICT r1 weak #1: the engine's free-running 5-level image sits ~4x further from the fp64 oracle than the
reference's own fp32 arithmetic.  This script replays the ORACLE pipeline with the engine's number formats
emulated in NumPy/torch-CPU, one ingredient at a time:

  f32          plain fp32 (the reference's arithmetic)                       -> the "noise floor"
  f32_order    fp32, but every conv accumulates its 9 taps separately (another legal fp32 summation order)
  store22      fp32 convs, every activation and weight rounded to the SPF16 pair hi+lo (fp16+fp16)
  prod3        store22 + products hi*hi + hi*lo + lo*hi only (lo*lo dropped), exact accumulation
  *_ws         the same with every layer's weights scaled by a power of two (max|w| -> [512,1024)) before the split,
               so that the weights' lo plane stays out of the fp16 subnormal range (He-normal weights ~0.02: lo ~5e-6)

Every run is compared with the fp64 run of the same pipeline.  Usage: python tests/noise_split_cpu.py [size] [seeds]
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import nets  # noqa: E402
from wct_tf_b200.weights import make_synthetic_weights  # noqa: E402

ALL = ["relu5_1", "relu4_1", "relu3_1", "relu2_1", "relu1_1"]


def spf16(t):
    """value stored by the engine for fp32 t: fp16 hi + fp16 lo."""
    hi = t.to(torch.float16).to(torch.float32)
    lo = (t - hi).to(torch.float16).to(torch.float32)
    return hi, lo


def make_conv(mode):
    def conv(x, w_hwio, b, pad):
        w = torch.from_numpy(np.ascontiguousarray(np.transpose(w_hwio, (3, 2, 0, 1)))).to(x.dtype)
        bb = torch.from_numpy(np.asarray(b)).to(x.dtype)
        if pad:
            x = F.pad(x, (1, 1, 1, 1), mode="reflect")
        if mode == "f32" or x.dtype == torch.float64:
            return F.conv2d(x, w, bb)
        if mode == "f32_order":
            kh, kw = w.shape[2], w.shape[3]
            H, W = x.shape[2] - kh + 1, x.shape[3] - kw + 1
            acc = None
            for i in range(kh):
                for j in range(kw):
                    t = F.conv2d(x[:, :, i:i + H, j:j + W], w[:, :, i:i + 1, j:j + 1])
                    acc = t if acc is None else acc + t
            return acc + bb.view(1, -1, 1, 1)
        xh, xl = spf16(x)
        wh, wl = spf16(w)
        if mode.endswith("_ws"):      # per-layer power-of-two weight scale: keeps the lo plane out of the fp16 subnormals
            sc = 2.0 ** (10 - int(np.ceil(np.log2(float(w.abs().max())))))
            wh, wl = spf16(w * sc)
            wh, wl = wh / sc, wl / sc
            mode_ = mode[:-3]
        else:
            mode_ = mode
        if mode_ == "store22":
            y = F.conv2d(xh + xl, wh + wl, bb)
        elif mode_ == "prod3":
            d = torch.float64
            y = (F.conv2d(xh.to(d), wh.to(d)) + F.conv2d(xh.to(d), wl.to(d)) + F.conv2d(xl.to(d), wh.to(d))).to(torch.float32)
            y = y + bb.view(1, -1, 1, 1)
        else:
            raise ValueError(mode)
        return y
    return conv


def run(mode, content, style, weights, dtype):
    old = nets._conv
    nets._conv = make_conv(mode)
    try:
        return nets.pipeline(content, style, weights, ALL, alpha=0.8, semantics="np", dtype=dtype)
    finally:
        nets._conv = old


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    weights = make_synthetic_weights(42)
    modes = ["f32", "f32_order", "store22", "prod3", "store22_ws", "prod3_ws"]
    print("free-running 5-level max-abs error vs the fp64 run, %dx%d, alpha 0.8, wct_np semantics" % (size, size))
    print("%-6s " % "seed" + " ".join("%-10s" % m for m in modes))
    rows = []
    for sd in range(seeds):
        rng = np.random.default_rng(1000 + sd)
        c = rng.integers(0, 256, (size, size, 3), dtype=np.uint8)
        s = rng.integers(0, 256, (size, size, 3), dtype=np.uint8)
        ref = run("f32", c, s, weights, np.float64)
        errs = [float(np.abs(run(m, c, s, weights, np.float32) - ref).max()) for m in modes]
        rows.append(errs)
        print("%-6d " % sd + " ".join("%-10.2e" % e for e in errs))
    rows = np.array(rows)
    print("%-6s " % "gmean" + " ".join("%-10.2e" % np.exp(np.log(rows[:, i]).mean()) for i in range(len(modes))))


if __name__ == "__main__":
    main()
