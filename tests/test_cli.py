"""CLI surface: the reference's flags (stylize.py:16-37) are all accepted with the same
defaults; image helpers behave like utils.py."""
import numpy as np

import stylize
from wct_tf_b200 import imageio as io


def test_reference_flags_and_defaults():
    p = stylize.build_parser()
    a = p.parse_args(["--relu-targets", "relu5_1", "relu1_1", "--checkpoints", "a", "b"])
    assert a.relu_targets == ["relu5_1", "relu1_1"] and a.checkpoints == ["a", "b"]
    assert (a.alpha, a.passes, a.device, a.adain, a.keep_colors, a.concat, a.swap5) == (1, 1, "/gpu:0", False, False, False, False)
    assert (a.style_size, a.crop_size, a.content_size, a.random) == (0, 0, 0, 0)
    assert (a.ss_alpha, a.ss_patch_size, a.ss_stride) == (0.6, 3, 1)
    b = p.parse_args(["--relu-targets", "relu1_1", "--alpha", "0.8", "--adain", "--style-size", "512", "-r", "2"])
    assert b.alpha == 0.8 and b.adain and b.style_size == 512 and b.random == 2


def test_resize_and_crop_semantics():
    img = np.random.default_rng(0).integers(0, 256, (40, 60, 3), dtype=np.uint8)
    r = io.resize_to(img, 20)
    assert r.shape == (20, 30, 3)                     # short side -> 20 (utils.py:55-67)
    r2 = io.resize_to(np.transpose(img, (1, 0, 2)), 20)
    assert r2.shape == (30, 20, 3)
    c = io.center_crop(img, 32)
    assert c.shape == (32, 32, 3) and np.array_equal(c, img[4:36, 14:46])
    up = io.center_crop(img[:10, :10], 16)            # too small -> upscale first (utils.py:32-34)
    assert up.shape == (16, 16, 3)


def test_coral_matches_target_statistics():
    rng = np.random.default_rng(1)
    src = rng.random((16, 16, 3)) * 0.5
    tgt = rng.random((12, 20, 3)) * np.array([0.2, 0.9, 0.4]) + 0.1
    out = io.coral(src, tgt)
    assert out.shape == src.shape
    assert np.allclose(out.reshape(-1, 3).mean(0), tgt.reshape(-1, 3).mean(0), atol=1e-6)
    u8 = io.preserve_colors_np(np.uint8(src * 255), np.uint8(tgt * 255))
    assert u8.dtype == np.uint8 and u8.shape == src.shape


def test_video_driver_batches_frames_in_order(tmp_path):
    """stylize_video.py: frames are numerically ordered, batched with ONE shared style, passes/concat are applied like
    stylize_video.py:114-130, ragged frames fall back to per-frame predict; no GPU needed (fake WCT)."""
    import numpy as np
    from PIL import Image
    import stylize_video as V

    frames = tmp_path / "clip"
    frames.mkdir()
    for i in [1, 2, 10, 11, 3]:                                 # lexicographic order would put 10, 11 before 2
        Image.fromarray(np.full((8, 12, 3), i, dtype=np.uint8)).save(str(frames / ("frame_%d.png" % i)))
    Image.fromarray(np.full((16, 12, 3), 4, dtype=np.uint8)).save(str(frames / "frame_4.png"))   # a ragged frame
    style = tmp_path / "style.png"
    Image.fromarray(np.full((6, 6, 3), 200, dtype=np.uint8)).save(str(style))
    calls = []

    class FakeWCT(object):
        def predict_batch(self, contents, styles, alpha=1, adain=False, passes=None, swap5=False, **kw):
            contents = np.asarray(contents)
            if passes is None:                                      # the per-frame path: one pass per call
                contents = contents[None] if contents.ndim == 3 else contents
                calls.append(("single", int(contents[0, 0, 0, 0])))
                return contents + 1
            calls.append(("batch", [int(c[0, 0, 0]) for c in contents], styles.shape[0], passes))
            return contents + passes

    class HostImageOps(object):
        """stand-in for wct_tf_b200.device_image (which needs a GPU): the frame logic only moves arrays around"""
        to_device = staticmethod(lambda img, device=None: np.asarray(img))
        to_host = staticmethod(lambda img: np.asarray(img))
        resize_to = staticmethod(lambda img, size: img)

        @staticmethod
        def concat_with_style(style, result):
            from wct_tf_b200 import imageio
            side = result.shape[0]
            return np.hstack([imageio._imresize(style, (side, side)), result])

    out = tmp_path / "out"
    n = V.main(["--relu-targets", "relu1_1", "--in-path", str(frames), "--style-path", str(style), "--out-path", str(out),
                "--batch", "2", "--passes", "2", "--concat"], wct_factory=lambda a: FakeWCT(), image_ops=HostImageOps)
    assert n == 6
    assert calls[0] == ("batch", [1, 2], 1, 2)                                # first batch: one shared style, both passes in one call
    assert [c[1] for c in calls if c[0] == "batch"] == [[1, 2], [10, 11]]     # frames 3|4 differ in size -> per-frame
    assert [c[1] for c in calls if c[0] == "single"] == [3, 4, 4, 5]          # 3 (two passes), then the ragged 4 (two passes)
    res = np.array(Image.open(str(out / "clip_style_frames" / "frame_10.png")))
    assert res.shape == (8, 8 + 12, 3) and res[0, -1, 0] == 12 and res[0, 0, 0] == 200   # --concat: [style | stylised]
    assert V.frame_key("frame_12.png") > V.frame_key("frame_2.png")


def test_keep_colors_matches_reference_golden():
    """--keep-colors = utils.preserve_colors_np -> coral.coral_numpy of the reference (tests/golden/make_golden.py), including
    its non-symmetric ``matSqrt`` (coral.py:8-11)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "coral_keep_colors.npz"))
    c = io.coral(g["style"] / 255., g["content"] / 255.)
    assert np.abs(c - g["coraled"]).max() < 1e-9
    assert np.array_equal(io.preserve_colors_np(g["style"], g["content"]), g["out"])
