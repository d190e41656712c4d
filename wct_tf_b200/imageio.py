"""Host-side image helpers of the CLI (restating utils.py:11-67, coral.py:8-39 with
PIL/NumPy: ``scipy.misc.imread/imresize/imsave`` no longer exist).  Off the hot path."""
from __future__ import annotations

import os

import numpy as np
from PIL import Image


def get_files(img_dir):
    """utils.py:11-17"""
    return [os.path.join(img_dir, x) for x in os.listdir(img_dir)]


def get_img(src):
    """utils.py:23-27: RGB uint8 HxWx3"""
    return np.asarray(Image.open(src).convert("RGB"))


def save_img(out_path, img):
    """utils.py:19-21"""
    Image.fromarray(np.clip(img, 0, 255).astype(np.uint8)).save(out_path)


def _imresize(img, hw):
    # scipy.misc.imresize(interp='bilinear') was PIL's bilinear resize on the uint8 image
    return np.asarray(Image.fromarray(np.asarray(img, dtype=np.uint8)).resize((int(hw[1]), int(hw[0])), Image.BILINEAR))


def resize_to(img, resize=512):
    """utils.py:55-67: resize the SHORT side to ``resize`` keeping the aspect ratio."""
    height, width = img.shape[0], img.shape[1]
    if height < width:
        shape = (resize, int(round(width / (height / resize))))
    else:
        shape = (int(round(height / (width / resize))), resize)
    return _imresize(img, shape)


def center_crop(img, size=256):
    """utils.py:29-38"""
    height, width = img.shape[0], img.shape[1]
    if height < size or width < size:
        img = resize_to(img, resize=size)
        height, width = img.shape[0], img.shape[1]
    h_off, w_off = (height - size) // 2, (width - size) // 2
    return img[h_off:h_off + size, w_off:w_off + size]


def _mat_sqrt(x):
    """coral.py:8-11 verbatim in effect: ``U, D, V = np.linalg.svd(x); U diag(sqrt D) V.T``.  numpy returns V^H as the third
    value, so for the symmetric input this is U sqrt(D) U -- NOT the symmetric square root U sqrt(D) U^T.  The reference's
    --keep-colors output depends on it, so it is reproduced (pinned by tests/golden/coral_keep_colors.npz)."""
    u, d, vh = np.linalg.svd(x)
    return (u * np.sqrt(d)) @ vh.T


def coral(source, target):
    """CORAL colour alignment of ``source`` to ``target`` statistics (coral.py:13-39):
    per-channel standardise, match the (cov + I) of the 3x3 channel covariance."""
    c = source.shape[-1]
    s = source.reshape(-1, c).T
    t = target.reshape(-1, c).T
    sm, ss = s.mean(1, keepdims=True), s.std(1, keepdims=True)
    tm, ts = t.mean(1, keepdims=True), t.std(1, keepdims=True)
    sn, tn = (s - sm) / ss, (t - tm) / ts
    cs = sn @ sn.T + np.eye(c)
    ct = tn @ tn.T + np.eye(c)
    out = _mat_sqrt(ct) @ np.linalg.inv(_mat_sqrt(cs)) @ sn
    out = out * ts + tm
    return out.T.reshape(source.shape)


def preserve_colors_np(style_rgb, content_rgb):
    """utils.py:87-90 (--keep-colors)"""
    return np.uint8(np.clip(coral(style_rgb / 255., content_rgb / 255.), 0, 1) * 255.)
