"""Timing of the device image steps (scope row 8f-3) next to Pillow / NumPy on the host: 1080p -> short side 512 resize,
512x512 centre crop, CORAL 512x512.  HBM-bound byte work; reported as us per image and GB/s of compulsory traffic."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from PIL import Image
from wct_tf_b200 import device_image as D
from oracle import image_ops as O


def gpu_time(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def cpu_time(fn, n=5):
    fn()
    t = time.time()
    for _ in range(n):
        fn()
    return (time.time() - t) / n * 1e6


rng = np.random.default_rng(0)
for (h, w, n) in [(1080, 1920, 1), (1080, 1920, 16), (512, 512, 16)]:
    img = rng.integers(0, 256, (n, h, w, 3), dtype=np.uint8)
    d = torch.from_numpy(img).cuda()
    oh, ow = (512, 910) if h == 1080 else (1024, 1024)
    us = gpu_time(lambda: D.imresize(d, (oh, ow)))
    cpu = cpu_time(lambda: [np.asarray(Image.fromarray(im).resize((ow, oh), Image.BILINEAR)) for im in img])
    byts = n * 3 * (h * w + oh * ow)
    print("resize %dx%d -> %dx%d  batch %2d : %8.1f us/batch  %7.1f GB/s (read+write once)   Pillow on the host %9.1f us  (x%.0f)"
          % (h, w, oh, ow, n, us, byts / us / 1e3, cpu, cpu / us))
s, c = torch.from_numpy(rng.integers(0, 256, (512, 512, 3), dtype=np.uint8)).cuda(), torch.from_numpy(rng.integers(0, 256, (512, 512, 3), dtype=np.uint8)).cuda()
us = gpu_time(lambda: D.preserve_colors_np(s, c), 20)
cpu = cpu_time(lambda: O.preserve_colors(s.cpu().numpy(), c.cpu().numpy()))
print("keep-colors (CORAL) 512x512 : %8.1f us (2 moment launches + one 72-byte D2H each + 3x3 host algebra + apply)   NumPy on the host %9.1f us (x%.0f)" % (us, cpu, cpu / us))
