"""GPU parity for scope row 8f-3: resize / crop / CORAL / concat on the device, through the C-ABI, against the oracle
restatement (oracle/image_ops.py) AND against Pillow itself (the third-party code the reference's scipy.misc.imresize ran).
Integer/byte work: the bar is bit-exact; the double-precision CORAL map is allowed one uint8 LSB on a vanishing fraction of
pixels (the reference truncates clip(x)*255, so a 1e-16 relative difference in the 3x3 statistics can flip a byte)."""
import os

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import image_ops as O
from wct_tf_b200 import _capi
from wct_tf_b200 import device_image as D
from tests import gpu_util as U

pytestmark = pytest.mark.gpu


def _img(h, w, c=3, seed=0, n=None):
    shape = (h, w, c) if n is None else (n, h, w, c)
    return np.random.default_rng(seed).integers(0, 256, shape, dtype=np.uint8)


def _smooth(h, w, seed):
    """a natural-ish image (low-pass noise): resampling error patterns differ from white noise"""
    r = np.random.default_rng(seed).standard_normal((h // 8 + 2, w // 8 + 2, 3))
    big = np.kron(r, np.ones((8, 8, 1)))[:h, :w]
    return np.uint8(np.clip(128 + 60 * big, 0, 255))


RESIZE_CASES = [(37, 53, 20, 29), (37, 53, 74, 91), (64, 64, 64, 64), (100, 40, 512, 205), (513, 301, 256, 150), (5, 7, 1, 1),
                (1, 1, 9, 4), (720, 1280, 512, 910), (1080, 1920, 512, 910), (33, 70, 33, 35), (33, 70, 66, 70),
                (300, 400, 299, 401), (2, 3, 512, 512), (512, 512, 1024, 1024)]


@pytest.mark.parametrize("h,w,oh,ow", RESIZE_CASES)
def test_resize_is_pillow_bit_for_bit(h, w, oh, ow):
    img = _img(h, w, seed=h * 7 + w)
    got = D.imresize(U.dev(img), (oh, ow)).cpu().numpy()
    want = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR))
    assert got.shape == want.shape and np.array_equal(got, want)
    if h * w <= 400 * 400:
        assert np.array_equal(got, O.resample_bilinear_u8(img, oh, ow))


@pytest.mark.parametrize("c", [1, 3, 4])
def test_resize_batches_channels_and_windows(c):
    """N > 1, 1 / 3 / 4 channels, and the crop window of utils.center_crop folded into the launch (only the window is
    computed): every frame equals the oracle's full resample, cropped."""
    imgs = _img(45, 61, c=c, seed=c, n=3)
    full = np.stack([O.resample_bilinear_u8(im, 70, 33) for im in imgs])
    got = D.imresize(U.dev(imgs), (70, 33)).cpu().numpy()
    assert np.array_equal(got, full)
    for (y0, x0, ho, wo) in [(0, 0, 70, 33), (5, 7, 40, 20), (69, 32, 1, 1), (0, 3, 70, 11)]:
        win = D.imresize(U.dev(imgs), (70, 33), window=(y0, x0, ho, wo)).cpu().numpy()
        assert np.array_equal(win, full[:, y0:y0 + ho, x0:x0 + wo])
    # one axis unchanged: Pillow skips that pass (Resample.c ImagingResampleInner); plain crop when both are unchanged
    for (oh, ow) in [(45, 30), (90, 61), (45, 61)]:
        want = np.stack([O.resample_bilinear_u8(im, oh, ow) for im in imgs])[:, 2:40, 1:29]
        assert np.array_equal(D.imresize(U.dev(imgs), (oh, ow), window=(2, 1, 38, 28)).cpu().numpy(), want)


@pytest.mark.parametrize("h,w", [(40, 64), (64, 40), (37, 37), (300, 451)])
def test_resize_to_center_crop_center_crop_to_match_the_reference_helpers(h, w):
    """utils.resize_to / center_crop / center_crop_to (utils.py:29-67) incl. the upscale-when-too-small branches and the
    Python-3 banker's rounding of the long side."""
    img = _smooth(h, w, seed=h + w)
    d = U.dev(img)
    for size in [16, 25, 96]:
        assert np.array_equal(D.resize_to(d, size).cpu().numpy(), O.resize_to(img, size))
        assert np.array_equal(D.center_crop(d, size).cpu().numpy(), O.center_crop(img, size))
    for (ht, wt) in [(32, 32), (h, w), (h + 9, w), (h, 2 * w), (96, 112)]:
        assert np.array_equal(D.center_crop_to(d, ht, wt).cpu().numpy(), O.center_crop_to(img, ht, wt))


def test_rgb_moments_are_exact_integers():
    img = _img(257, 331, seed=5)
    npix, s1, s2 = D.rgb_moments(U.dev(img))
    x = img.reshape(-1, 3).astype(np.int64)
    assert npix == x.shape[0]
    assert [int(v) for v in s1] == list(x.sum(0))
    assert np.array_equal(np.array(s2, dtype=np.int64), x.T @ x)


def _lsb_report(got, want):
    d = np.abs(got.astype(int) - want.astype(int))
    return int(d.max()), float((d > 0).mean())


@pytest.mark.parametrize("hs,ws,hc,wc", [(48, 64, 40, 56), (300, 200, 128, 512), (512, 512, 512, 512)])
def test_keep_colors_matches_the_coral_restatement(hs, ws, hc, wc):
    """utils.preserve_colors_np -> coral.coral_numpy (incl. its non-symmetric matSqrt): uint8 result within one LSB on at most
    1e-4 of the bytes (measured: identical)."""
    style, content = _smooth(hs, ws, 1), _smooth(hc, wc, 2)
    got = D.preserve_colors_np(U.dev(style), U.dev(content)).cpu().numpy()
    want = O.preserve_colors(style, content)
    mx, frac = _lsb_report(got, want)
    assert got.shape == want.shape and mx <= 1 and frac <= 1e-4, (mx, frac)


def test_keep_colors_reference_golden():
    """the fixture written by the reference's own coral.py / utils.preserve_colors_np (tests/golden/make_golden.py)"""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "coral_keep_colors.npz"))
    got = D.preserve_colors_np(U.dev(g["style"]), U.dev(g["content"])).cpu().numpy()
    mx, frac = _lsb_report(got, g["out"])
    assert mx <= 1 and frac <= 1e-4, (mx, frac)


def test_concat_thumbnail():
    style, res = _img(30, 44, seed=1), _img(24, 40, seed=2)
    got = D.concat_with_style(U.dev(style), U.dev(res)).cpu().numpy()
    assert np.array_equal(got, np.hstack([O.resample_bilinear_u8(style, 24, 24), res]))      # stylize.py:107-111


def test_bad_arguments_are_reported_not_launched():
    lib = _capi.load()
    d = U.dev(_img(8, 8))
    out = torch.empty(64 * 3, dtype=torch.uint8, device="cuda")
    ws = torch.empty(1 << 16, dtype=torch.uint8, device="cuda")
    assert lib.wctb200_resize_workspace_bytes(1, 8, 8, 3, 0, 4, 4) == 0
    rc = lib.wctb200_resize_bilinear_u8(d.data_ptr(), 1, 8, 8, 3, 4, 4, 2, 2, 4, 4, out.data_ptr(), ws.data_ptr(), 1 << 16, U.stream())
    assert rc == _capi.EINVAL and b"window" in lib.wctb200_last_error()
    rc = lib.wctb200_resize_bilinear_u8(d.data_ptr(), 1, 8, 8, 3, 4, 4, 0, 0, 4, 4, out.data_ptr(), ws.data_ptr(), 16, U.stream())
    assert rc == _capi.EWS
    with pytest.raises(_capi.WctB200Error):
        D.imresize(torch.zeros(4, 4, 3, dtype=torch.uint8), (2, 2))          # host tensor: no CPU fallback


def test_cli_device_image_path_equals_host_composition(tmp_path):
    """stylize.py --content-size --style-size --crop-size --keep-colors --passes 2 --concat: the file written equals
    resize/crop/CORAL by the oracle -> WCT.predict twice -> concat, i.e. the device-resident flow changes nothing."""
    import stylize
    from wct_tf_b200.wct import WCT
    from wct_tf_b200.weights import make_synthetic_weights
    cdir, odir = tmp_path / "c", tmp_path / "o"
    cdir.mkdir()
    content, style = _smooth(90, 70, 11), _smooth(77, 120, 12)
    Image.fromarray(content).save(cdir / "a.png")
    Image.fromarray(style).save(tmp_path / "st.png")
    targets = ["relu2_1", "relu1_1"]
    stylize.main(["--synthetic-weights", "42", "--relu-targets"] + targets + ["--content-path", str(cdir), "--style-path",
                  str(tmp_path / "st.png"), "--out-path", str(odir), "--alpha", "0.7", "--content-size", "48", "--style-size", "40",
                  "--crop-size", "32", "--keep-colors", "--passes", "2", "--concat"])
    got = np.asarray(Image.open(odir / "a_st.png"))
    c = O.resize_to(content, 48)
    s = O.preserve_colors(O.center_crop(O.resize_to(style, 40), 32), c)
    wct = WCT(checkpoints=None, relu_targets=targets, vgg_path=None, weights=make_synthetic_weights(42, relu_targets=targets))
    r = wct.predict(wct.predict(c, s, alpha=0.7), s, alpha=0.7)
    want = np.hstack([O.resample_bilinear_u8(s, r.shape[0], r.shape[0]), r])
    assert got.shape == want.shape and np.array_equal(got, want)
