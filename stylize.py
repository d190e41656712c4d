#!/usr/bin/env python
"""Drop-in for the reference CLI (stylize.py:14-122): same flags, same content x style loop,
same output naming -- the TF session is replaced by the B200 engine.

  python stylize.py --checkpoints dec5.npz dec4.npz ... --relu-targets relu5_1 relu4_1 ... \
      --vgg-path vgg.npz --content-path IN --style-path STYLE --out-path OUT --alpha 0.8

Weight files are ``.npz`` bundles (wct_tf_b200.weights.save_weights); ``--synthetic-weights SEED``
runs with seeded random weights when the published models are not on disk (offline build).
"""
from __future__ import division, print_function

import argparse
import os
import time

import numpy as np


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument('--checkpoints', nargs='+', type=str, help='List of checkpoint files (one per relu target)')
    p.add_argument('--relu-targets', nargs='+', type=str, help='List of reluX_1 layers, corresponding to --checkpoints', required=True)
    p.add_argument('--vgg-path', type=str, help='Path to the normalised VGG19 weights', default='models/vgg_normalised.npz')
    p.add_argument('--content-path', type=str, dest='content_path', help='Content image or folder of images')
    p.add_argument('--style-path', type=str, dest='style_path', help='Style image or folder of images')
    p.add_argument('--out-path', type=str, dest='out_path', help='Output folder path')
    p.add_argument('--keep-colors', action='store_true', help="Preserve the colors of the style image", default=False)
    p.add_argument('--device', type=str, help='Device to perform compute on, e.g. /gpu:0', default='/gpu:0')
    p.add_argument('--style-size', type=int, help="Resize style image to this size before cropping", default=0)
    p.add_argument('--crop-size', type=int, help="Crop square size", default=0)
    p.add_argument('--content-size', type=int, help="Resize short side of content image to this", default=0)
    p.add_argument('--passes', type=int, help="# of stylization passes per content image", default=1)
    p.add_argument('-r', '--random', type=int, help="Choose # of random subset of images from style folder", default=0)
    p.add_argument('--alpha', type=float, help="Alpha blend value", default=1)
    p.add_argument('--concat', action='store_true', help="Concatenate style image and stylized output", default=False)
    p.add_argument('--adain', action='store_true', help="Use AdaIN instead of WCT", default=False)
    # style swap at relu5_1 (ops.py:145-278); built for --ss-patch-size 3 --ss-stride 1
    p.add_argument('--swap5', action='store_true', help="Swap style on layer relu5_1", default=False)
    p.add_argument('--ss-alpha', type=float, default=0.6)
    p.add_argument('--ss-patch-size', type=int, default=3)
    p.add_argument('--ss-stride', type=int, default=1)
    # additions
    p.add_argument('--synthetic-weights', type=int, default=None, help="use seeded random weights (no model files needed)")
    p.add_argument('--semantics', type=str, default='tf', choices=['tf', 'np'], help="wct_tf (reference graph) or wct_np blend/eps semantics")
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    from wct_tf_b200 import imageio as io
    from wct_tf_b200.wct import WCT
    start = time.time()
    weights = None
    if args.synthetic_weights is not None:
        from wct_tf_b200.weights import make_synthetic_weights
        weights = make_synthetic_weights(args.synthetic_weights, relu_targets=args.relu_targets)
    elif not args.checkpoints:
        raise SystemExit("--checkpoints is required (or --synthetic-weights SEED)")
    wct_model = WCT(checkpoints=args.checkpoints, relu_targets=args.relu_targets, vgg_path=args.vgg_path,
                    device=args.device, ss_patch_size=args.ss_patch_size, ss_stride=args.ss_stride,
                    weights=weights, semantics=args.semantics, verbose=True)

    content_files = io.get_files(args.content_path) if os.path.isdir(args.content_path) else [args.content_path]
    if os.path.isdir(args.style_path):
        style_files = io.get_files(args.style_path)
        if args.random > 0:
            style_files = np.random.choice(style_files, args.random)
    else:
        style_files = [args.style_path]
    os.makedirs(args.out_path, exist_ok=True)

    count = 0
    for content_fullpath in content_files:                       # stylize.py:70
        content_prefix, content_ext = os.path.splitext(content_fullpath)
        content_prefix = os.path.basename(content_prefix)
        content_img = io.get_img(content_fullpath)
        if args.content_size > 0:
            content_img = io.resize_to(content_img, args.content_size)
        for style_fullpath in style_files:                       # stylize.py:78
            style_prefix = os.path.basename(os.path.splitext(style_fullpath)[0])
            style_img = io.get_img(style_fullpath)
            if args.style_size > 0:
                style_img = io.resize_to(style_img, args.style_size)
            if args.crop_size > 0:
                style_img = io.center_crop(style_img, args.crop_size)
            if args.keep_colors:
                style_img = io.preserve_colors_np(style_img, content_img)
            stylized_rgb = wct_model.predict(content_img, style_img, args.alpha, args.swap5, args.ss_alpha, args.adain)
            for _ in range(args.passes - 1):                     # stylize.py:102-104
                stylized_rgb = wct_model.predict(stylized_rgb, style_img, args.alpha, args.swap5, args.ss_alpha, args.adain)
            if args.concat:                                      # stylize.py:107-111
                side = stylized_rgb.shape[0]
                style_resized = io._imresize(style_img, (side, side))
                stylized_rgb = np.hstack([style_resized, stylized_rgb])
            out_f = os.path.join(args.out_path, '{}_{}{}'.format(content_prefix, style_prefix, content_ext))
            io.save_img(out_f, stylized_rgb)
            count += 1
            print("{}: Wrote stylized output image to {}".format(count, out_f))
    print("Finished stylizing {} outputs in {}s".format(count, time.time() - start))


if __name__ == '__main__':
    main()
