"""encoder head (3 -> 64 @ HxW): tensor-core kernel vs the SIMT kernel.  python tools/head_bench.py [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from wct_tf_b200 import _capi
from tests import gpu_util as U

lib = _capi.load()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rng = np.random.default_rng(0)
for (h, w) in [(512, 512), (1024, 1024), (256, 256)]:
    n = N if h <= 512 else max(1, N // 4)
    img = torch.from_numpy(rng.random((n, h, w, 3)).astype(np.float32)).cuda()
    k = torch.from_numpy((rng.normal(0, 1, (27, 64)) * 0.2).astype(np.float32)).cuda()
    b = torch.from_numpy(rng.normal(0, 0.1, 64).astype(np.float32)).cuda()
    outs = {}
    for impl in (1, 0):
        lib.wctb200_debug_set_conv_head_tc(impl)
        out = torch.empty(lib.wctb200_act_bytes(n, h, w, 64), dtype=torch.uint8, device="cuda")
        f = lambda: _capi.check(lib.wctb200_conv_head(img.data_ptr(), n, h, w, k.data_ptr(), b.data_ptr(), out.data_ptr(), U.stream()))
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        byts = n * h * w * 12 + n * (h + 2) * (w + 2) * 64 * 4
        res = torch.empty((n, h, w, 64), dtype=torch.float32, device="cuda")
        _capi.check(lib.wctb200_act_to_f32(out.data_ptr(), n, h, w, 64, res.data_ptr(), U.stream()))
        outs[impl] = res
        print("%dx%d x%d  %-11s %7.3f ms  %6.0f GB/s compulsory (read image + write activation once)" % (h, w, n, "tensor-core" if impl else "simt", ms, byts / ms / 1e6))
    print("   max |tc - simt| = %.2e (values up to %.1f)" % ((outs[1] - outs[0]).abs().max().item(), outs[0].abs().max().item()))
lib.wctb200_debug_set_conv_head_tc(1)
