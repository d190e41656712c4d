#!/usr/bin/env python
"""Video / frame-sequence driver with the reference's flags (stylize_video.py:17-42 of eridgd/WCT-TF).

The reference splits the video into PNG frames with ffmpeg, calls ``WCT.predict`` once per frame and re-encodes
(stylize_video.py:75-149).  Here frames go through ``WCT.predict_batch`` in batches that share ONE style image (the
style side of every level is computed once per batch), while a thread pool decodes the next batch and writes the
previous one, so the GPU never waits for PNG I/O.  ``--in-path`` may also be a directory of frames (then no ffmpeg is
needed and the stylised frames are left in ``--out-path``); ffmpeg is used only when it is installed.
Additions: ``--batch``, ``--synthetic-weights``, ``--adain``.
"""
from __future__ import division, print_function

import argparse
import os
import random
import shutil
import subprocess
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np


# video-specific flags; the stylisation flags are shared with stylize.py (same names and defaults as the reference,
# stylize_video.py:17-42)
_VIDEO_FLAGS = [
    (("--in-path",), dict(type=str, required=True, help="video file, or a directory of frames")),
    (("--out-path",), dict(type=str, required=True, help="folder for the encoded videos / stylised frames")),
    (("--style-path",), dict(type=str, required=True, help="style image, or a folder of them")),
    (("--tmp-dir",), dict(type=str, dest="tmp_dir", default=None, help="where ffmpeg unpacks the frames")),
    (("--keep-tmp",), dict(action="store_true", default=False, help="keep the unpacked frames")),
    (("--batch",), dict(type=int, default=15, help="frames per predict_batch call (not in the reference)")),
    (("--fps",), dict(type=int, default=30, help="frame rate of the encoded video (not in the reference: fixed 30 there)")),
]
_SHARED = ("--checkpoints", "--relu-targets", "--vgg-path", "--keep-colors", "--style-size", "--crop-size", "--content-size",
           "--passes", "--device", "--alpha", "--concat", "--swap5", "--ss-alpha", "--ss-patch-size", "--ss-stride", "--adain",
           "--synthetic-weights")


def build_parser():
    import stylize
    parser = argparse.ArgumentParser(description="stylise a video or a frame sequence")
    for names, kw in stylize._FLAGS:
        if names[-1] in _SHARED:
            parser.add_argument(*names, **kw)
    for names, kw in _VIDEO_FLAGS:
        parser.add_argument(*names, **kw)
    return parser


def frame_key(name):
    """frame_12.png sorts after frame_2.png (the reference iterates os.listdir order; ffmpeg numbers the frames)."""
    digits = ''.join(ch if ch.isdigit() else ' ' for ch in os.path.basename(name)).split()
    return (int(digits[-1]) if digits else -1, name)


def batches(seq, n):
    for i in range(0, len(seq), n):
        yield seq[i:i + n]


def stylize_frames(wct_model, in_files, out_files, style_img, args, io, pool, dimg, device):
    """All frames of one clip with one style (a uint8 CUDA tensor, already resized / cropped).  Frames of equal size are
    batched; the decode of batch i+1 and the PNG writes of batch i-1 overlap the GPU work of batch i.  Resize
    (``--content-size``), CORAL (``--keep-colors``), every pass and the ``--concat`` thumbnail run on the device
    (wct_tf_b200.device_image); a frame crosses PCIe once in each direction."""
    def finish(out_f, frame_u8):
        io.save_img(out_f, frame_u8)

    def prep(frames_dev):
        return dimg.resize_to(frames_dev, args.content_size) if args.content_size > 0 else frames_dev

    todo = list(batches(list(zip(in_files, out_files)), max(1, args.batch)))
    pending_loads = [pool.submit(io.get_img, f) for f, _ in todo[0]] if todo else []
    writes, count = [], 0
    for bi, group in enumerate(todo):
        frames = [f.result() for f in pending_loads]
        pending_loads = [pool.submit(io.get_img, f) for f, _ in todo[bi + 1]] if bi + 1 < len(todo) else []
        same = all(fr.shape == frames[0].shape for fr in frames) and not args.keep_colors and not args.swap5
        if same and len(frames) > 1:
            x = prep(dimg.to_device(np.stack(frames), device))
            # --passes (stylize_video.py:119-121) run back to back on the device
            out = wct_model.predict_batch(x, style_img[None], alpha=args.alpha, adain=args.adain, passes=args.passes,
                                          return_device=True)
            results = [(out[i], style_img) for i in range(len(frames))]
        else:                                                  # per-frame styles (CORAL), style swap or ragged sizes
            results = []
            for fr in frames:
                fr = prep(dimg.to_device(fr, device))
                style_rgb = dimg.preserve_colors_np(style_img, fr) if args.keep_colors else style_img
                kw = dict(alpha=args.alpha, ss_alpha=args.ss_alpha, adain=args.adain, return_device=True)
                o = wct_model.predict_batch(fr, style_rgb, swap5=args.swap5, **kw)
                for _ in range(args.passes - 1):
                    o = wct_model.predict_batch(o, style_rgb, swap5=False, **kw)
                results.append((o[0], style_rgb))
        for (_, out_f), (o, srgb) in zip(group, results):
            if args.concat:                                    # stylize_video.py:125-128
                o = dimg.concat_with_style(srgb, o)
            writes.append(pool.submit(finish, out_f, dimg.to_host(o)))
            count += 1
    for w in writes:
        w.result()
    return count


def have_ffmpeg():
    return shutil.which('ffmpeg') is not None


def main(argv=None, wct_factory=None, image_ops=None):
    """``wct_factory`` / ``image_ops`` let the tests drive the frame logic with stand-ins for the engine and for
    wct_tf_b200.device_image; the CLI always uses the real ones."""
    args = build_parser().parse_args(argv)
    from wct_tf_b200 import imageio as io
    if image_ops is None:
        from wct_tf_b200 import device_image as image_ops
    dimg = image_ops
    start = time.time()
    if wct_factory is None:
        from wct_tf_b200.wct import WCT
        weights = None
        if args.synthetic_weights is not None:
            from wct_tf_b200.weights import make_synthetic_weights
            weights = make_synthetic_weights(args.synthetic_weights, relu_targets=args.relu_targets)
        elif not args.checkpoints:
            raise SystemExit("--checkpoints is required (or --synthetic-weights SEED)")
        wct_model = WCT(checkpoints=args.checkpoints, relu_targets=args.relu_targets, vgg_path=args.vgg_path,
                        device=args.device, ss_patch_size=args.ss_patch_size, ss_stride=args.ss_stride, weights=weights)
    else:
        wct_model = wct_factory(args)

    style_files = io.get_files(args.style_path) if os.path.isdir(args.style_path) else [args.style_path]
    os.makedirs(args.out_path, exist_ok=True)
    from_dir = os.path.isdir(args.in_path)
    tmp_dir = args.tmp_dir or os.path.join(args.out_path, '_____fns_frames_%s' % random.randint(0, 99999))
    if from_dir:
        in_files = sorted(io.get_files(args.in_path), key=frame_key)
    else:
        if not have_ffmpeg():
            raise SystemExit("ffmpeg is not installed: pass a directory of frames as --in-path")
        in_dir = os.path.join(tmp_dir, 'input')
        os.makedirs(in_dir, exist_ok=True)
        subprocess.check_call(['ffmpeg', '-i', args.in_path, os.path.join(in_dir, 'frame_%d.png')])   # stylize_video.py:75-81
        in_files = sorted([os.path.join(in_dir, x) for x in os.listdir(in_dir)], key=frame_key)
    clip = os.path.basename(os.path.normpath(os.path.splitext(args.in_path)[0] if not from_dir else args.in_path))
    ext = '.mp4' if from_dir else os.path.splitext(args.in_path)[1]

    total = 0
    with ThreadPoolExecutor(max_workers=8) as pool:
        for style_fullpath in style_files:                     # stylize_video.py:97
            device = wct_model.engine.device if hasattr(wct_model, "engine") else None
            style_img = dimg.to_device(io.get_img(style_fullpath), device)
            if args.style_size > 0:
                style_img = dimg.resize_to(style_img, args.style_size)
            if args.crop_size > 0:
                style_img = dimg.center_crop(style_img, args.crop_size)
            style_prefix = os.path.basename(os.path.splitext(style_fullpath)[0])
            out_v = os.path.join(args.out_path, '{}_{}{}'.format(clip, style_prefix, ext))
            frames_dir = os.path.join(args.out_path if from_dir else tmp_dir, '{}_{}_frames'.format(clip, style_prefix))
            if os.path.isfile(out_v):                          # stylize_video.py:108-110
                print("SKIP", out_v)
                continue
            os.makedirs(frames_dir, exist_ok=True)
            out_files = [os.path.join(frames_dir, os.path.splitext(os.path.basename(f))[0] + '.png') for f in in_files]
            n = stylize_frames(wct_model, in_files, out_files, style_img, args, io, pool, dimg, device)
            total += n
            print("Stylized {} frames with {} -> {}".format(n, style_prefix, frames_dir))
            # stylize_video.py:137-149 re-encodes "frame_%d.png".  Frames extracted by ffmpeg carry that name; a user-supplied
            # directory of frames keeps its own basenames, so it is only re-encoded when they follow the same pattern
            named_like_ffmpeg = all(os.path.basename(f) == 'frame_%d.png' % (i + 1) for i, f in enumerate(out_files))
            if have_ffmpeg() and named_like_ffmpeg:
                pattern = os.path.join(frames_dir, 'frame_%d.png')
                subprocess.check_call(['ffmpeg', '-y', '-i', pattern, '-f', 'mp4', '-q:v', '0', '-vcodec', 'mpeg4',
                                       '-r', str(args.fps), out_v])
                print('Video at: %s' % out_v)
    if not from_dir and not args.keep_tmp and len(style_files) == 1:
        shutil.rmtree(tmp_dir, ignore_errors=True)
    dt = time.time() - start
    print('Processed {} frames in: {:.2f}s ({:.1f} frames/s incl. PNG I/O)'.format(total, dt, total / max(dt, 1e-9)))
    return total


if __name__ == '__main__':
    main()
