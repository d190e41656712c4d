"""Per-layer conv micro-benchmark: every conv shape of the 512x512 pipeline (3x3 layers and the UP2 layers that absorbed an
UpSampling2D), default dispatch vs fused MMA off.  TFLOP/s = EXECUTED algorithmic flops (2*taps*Cin*Cout per output pixel)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wct_tf_b200 import _capi
from tests import gpu_util as U

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
# (output H=W, Cin, Cout, up2)
SHAPES = [(512, 64, 64, 0), (256, 64, 128, 0), (256, 128, 128, 0), (256, 128, 64, 0), (128, 128, 256, 0), (128, 256, 256, 0),
          (128, 256, 128, 0), (64, 256, 512, 0), (64, 512, 512, 0), (64, 512, 256, 0), (32, 512, 512, 0),
          (64, 512, 512, 1), (128, 256, 256, 1), (256, 128, 128, 1), (512, 64, 64, 1)]
CFGS = [("default", -1), ("nofuse", 0), ("fuse", 1)]
lib = U.lib()
print("%-24s" % "HxW Cin->Cout" + "".join("%18s" % c[0] for c in CFGS) + "   (executed TFLOP/s, batch %d)" % B)
for hw, cin, cout, up2 in SHAPES:
    hin = hw // 2 if up2 else hw
    x = torch.randn(B, hin, hin, cin, device="cuda").relu_()
    xin = torch.empty(lib.wctb200_act_bytes(B, hin, hin, cin), dtype=torch.uint8, device="cuda")
    _capi.check(lib.wctb200_act_from_f32(x.data_ptr(), B, hin, hin, cin, xin.data_ptr(), U.stream()))
    k = (torch.randn(3, 3, cin, cout, device="cuda") * (2.0 / (9 * cin)) ** 0.5).contiguous()
    bias = torch.zeros(cout, device="cuda")
    ws = torch.empty(lib.wctb200_conv_weight_bytes(16 if up2 else 9, cin, cout), dtype=torch.uint8, device="cuda")
    if up2:
        _capi.check(lib.wctb200_prep_conv_weights_up2(k.data_ptr(), cin, cout, ws.data_ptr(), U.stream()))
    else:
        _capi.check(lib.wctb200_prep_conv_weights(k.data_ptr(), 9, cin, cout, ws.data_ptr(), U.stream()))
    out = torch.empty(lib.wctb200_act_bytes(B, hw, hw, cout), dtype=torch.uint8, device="cuda")
    flops = 2.0 * (4 if up2 else 9) * cin * cout * B * hw * hw
    row = "%-24s" % ("%dx%d %d->%d%s" % (hw, hw, cin, cout, " up2" if up2 else ""))
    for name, fuse in CFGS:
        lib.wctb200_debug_set_conv_fuse(fuse)
        fn = lib.wctb200_conv3x3_up2 if up2 else lib.wctb200_conv3x3
        def run():
            _capi.check(fn(xin.data_ptr(), B, hin, hin, cin, ws.data_ptr(), bias.data_ptr(), cout, 1, out.data_ptr(), U.stream()))
        for _ in range(3): run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        row += "%10.0f (%5.0fus)" % (flops / ms / 1e9, ms * 1e3)
    print(row, flush=True)
lib.wctb200_debug_set_conv_fuse(-1)
_capi.check(lib.wctb200_check_device(U.stream()))
