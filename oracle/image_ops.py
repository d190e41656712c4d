"""CPU oracle for the image pre/post steps of the CLI (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

Scope row 8f-3: ``utils.resize_to`` / ``center_crop`` / ``center_crop_to`` (utils.py:29-67), ``utils.preserve_colors_np`` ->
``coral.coral_numpy`` (utils.py:87-90, coral.py:8-39) and the ``--concat`` thumbnail (stylize.py:107-111).

Third-party arithmetic.  The reference resizes with ``scipy.misc.imresize(img, shape, interp='bilinear')`` (utils.py:48,67),
which built a PIL image from the uint8 array and called ``Image.resize(..., BILINEAR)``; neither scipy<1.3 nor the Pillow of
2017 is under /root/reference (requirements.txt pins neither).  The algorithm is Pillow's ``ImagingResample`` 8-bit path
(src/libImaging/Resample.c, unchanged in substance since Pillow 3.4): a separable triangle filter whose support grows with
the down-scaling factor, coefficients computed in double precision, normalised, converted to 22-bit fixed point, and applied
HORIZONTALLY FIRST with a rounding to uint8 between the two passes.  ``resample_bilinear_u8`` restates it; it is pinned
bit-for-bit against the Pillow installed in this image (12.2.0) by tests/test_oracle.py over up-scaling, down-scaling,
identity and one-pixel cases.
"""
from __future__ import annotations

import numpy as np

PRECISION_BITS = 32 - 8 - 2      # Resample.c: 8 bits of pixel, 2 bits of head-room for the accumulation


def resample_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the BILINEAR filter (support 1.0) over the box
    [0, in_size): returns (bounds [out_size, 2] = (xmin, count), kk [out_size, ksize] int32)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    bounds = np.zeros((out_size, 2), np.int64)
    kk = np.zeros((out_size, ksize), np.int32)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)            # C cast: truncation towards zero
        xmin = max(xmin, 0)
        xmax = int(center + support + 0.5)
        xmax = min(xmax, in_size) - xmin
        x = np.arange(xmax, dtype=np.float64)
        arg = (x + xmin - center + 0.5) * ss
        w = np.maximum(1.0 - np.abs(arg), 0.0)         # bilinear_filter
        ww = 0.0
        for v in w:                                    # sequential accumulation, as the C loop
            ww += v
        if ww != 0.0:
            w = w / ww
        kf = w * float(1 << PRECISION_BITS)
        kk[xx, :xmax] = np.where(kf < 0, (-0.5 + kf), (0.5 + kf)).astype(np.int64)   # (int) cast truncates
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img, out_size, axis):
    img = np.moveaxis(img, axis, 0).astype(np.int64)
    bounds, kk = resample_coeffs(img.shape[0], out_size)
    out = np.empty((out_size,) + img.shape[1:], np.uint8)
    for xx in range(out_size):
        x0, n = bounds[xx]
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for t in range(n):
            acc += img[x0 + t] * int(kk[xx, t])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return np.moveaxis(out, 0, axis)


def resample_bilinear_u8(img, out_h, out_w):
    """ImagingResample (8 bits per channel): horizontal pass when the width changes, then vertical pass when the height
    changes, each rounding to uint8 (Resample.c ImagingResampleInner)."""
    img = np.asarray(img, np.uint8)
    if img.shape[1] != out_w:
        img = _pass(img, out_w, 1)
    if img.shape[0] != out_h:
        img = _pass(img, out_h, 0)
    return img


def resize_to(img, resize=512):
    """utils.py:55-67"""
    h, w = img.shape[0], img.shape[1]
    if h < w:
        shape = (resize, int(round(w / (h / resize))))
    else:
        shape = (int(round(h / (w / resize))), resize)
    return resample_bilinear_u8(img, shape[0], shape[1])


def center_crop(img, size=256):
    """utils.py:29-38"""
    h, w = img.shape[0], img.shape[1]
    if h < size or w < size:
        img = resize_to(img, size)
        h, w = img.shape[0], img.shape[1]
    ho, wo = (h - size) // 2, (w - size) // 2
    return img[ho:ho + size, wo:wo + size]


def center_crop_to(img, H_target, W_target):
    """utils.py:40-53 (``imresize(img, <float>)`` scales both sides by the fraction and truncates)"""
    h, w = img.shape[0], img.shape[1]
    if h < H_target or w < W_target:
        rat = max(H_target / h, W_target / w)
        img = resample_bilinear_u8(img, int(h * rat), int(w * rat))
        h, w = img.shape[0], img.shape[1]
    ho, wo = (h - H_target) // 2, (w - W_target) // 2
    return img[ho:ho + H_target, wo:wo + W_target]


def mat_sqrt(x):
    """coral.py:8-11: ``U, D, V = np.linalg.svd(x); U * diag(sqrt(D)) * V.T``.  numpy's third return value is already V^T,
    so for the symmetric input this evaluates U sqrt(D) U, not the symmetric root; reproduced as written."""
    u, d, vh = np.linalg.svd(x)
    return (u * np.sqrt(d)) @ vh.T


def coral(source, target):
    """coral.py:13-39 on HxWxC float images in [0, 1]."""
    c = source.shape[-1]
    s = source.reshape(-1, c).T
    t = target.reshape(-1, c).T
    sm, ss = s.mean(1, keepdims=True), s.std(1, keepdims=True)
    tm, ts = t.mean(1, keepdims=True), t.std(1, keepdims=True)
    sn, tn = (s - sm) / ss, (t - tm) / ts
    cs = sn @ sn.T + np.eye(c)
    ct = tn @ tn.T + np.eye(c)
    out = mat_sqrt(ct) @ np.linalg.inv(mat_sqrt(cs)) @ sn
    out = out * ts + tm
    return out.T.reshape(source.shape)


def preserve_colors(style_rgb, content_rgb):
    """utils.py:87-90"""
    return np.uint8(np.clip(coral(style_rgb / 255., content_rgb / 255.), 0, 1) * 255.)
