"""Per-layer conv micro-benchmark: every 3x3 layer shape of the 512x512 pipeline, v2 vs v4."""
import os, sys, itertools
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wct_tf_b200 import _capi
from tests import gpu_util as U

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
SHAPES = [(512, 64, 64), (512, 64, 128), (256, 128, 128), (256, 128, 64), (256, 64, 128), (128, 128, 256), (128, 256, 256),
          (128, 256, 128), (64, 256, 512), (64, 512, 512), (64, 512, 256), (32, 512, 512)]
CFGS = [("v2", 5, 0, 0, -1), ("v2 nofuse", 5, 0, 0, 0), ("pairs", 6, 0, 0, -1), ("pairs bn64", 6, 0, 64, -1), ("v4 c2", 4, 2, 0, -1)]
if os.environ.get("CONV_BENCH_ONLY"):
    idx = [int(v) for v in os.environ["CONV_BENCH_ONLY"].split(",")]
    SHAPES = [SHAPES[i] for i in idx]
    CFGS = [CFGS[1]]
lib = U.lib()
rng = np.random.default_rng(0)
print("%-18s" % "HxW Cin->Cout" + "".join("%16s" % c[0] for c in CFGS) + "   (algorithmic TFLOP/s, batch %d)" % B)
for hw, cin, cout in SHAPES:
    x = torch.randn(B, hw, hw, cin, device="cuda").relu_()
    xin = torch.empty(lib.wctb200_act_bytes(B, hw, hw, cin), dtype=torch.uint8, device="cuda")
    _capi.check(lib.wctb200_act_from_f32(x.data_ptr(), B, hw, hw, cin, xin.data_ptr(), U.stream()))
    k = (torch.randn(3, 3, cin, cout, device="cuda") * (2.0 / (9 * cin)) ** 0.5).contiguous()
    bias = torch.zeros(cout, device="cuda")
    ws = torch.empty(lib.wctb200_conv_weight_bytes(9, cin, cout), dtype=torch.uint8, device="cuda")
    _capi.check(lib.wctb200_prep_conv_weights(k.data_ptr(), 9, cin, cout, ws.data_ptr(), U.stream()))
    out = torch.empty(lib.wctb200_act_bytes(B, hw, hw, cout), dtype=torch.uint8, device="cuda")
    flops = 2.0 * 9 * cin * cout * B * hw * hw
    row = "%-18s" % ("%dx%d %d->%d" % (hw, hw, cin, cout))
    for name, impl, cl, bn, fuse in CFGS:
        lib.wctb200_debug_set_conv_fuse(fuse)
        lib.wctb200_debug_set_conv_impl(impl)
        lib.wctb200_debug_set_conv4(cl if cl else 2, -1)
        lib.wctb200_debug_set_conv_bn(bn)
        def run():
            _capi.check(lib.wctb200_conv3x3(xin.data_ptr(), B, hw, hw, cin, ws.data_ptr(), bias.data_ptr(), cout, 1, out.data_ptr(), U.stream()))
        for _ in range(3): run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        row += "%9.0f (%4.0fus)" % (flops / ms / 1e9, ms * 1e3)
    print(row, flush=True)
lib.wctb200_debug_set_conv_fuse(-1); lib.wctb200_debug_set_conv_impl(2); lib.wctb200_debug_set_conv4(2, -1); lib.wctb200_debug_set_conv_bn(0)
_capi.check(lib.wctb200_check_device(U.stream()))
