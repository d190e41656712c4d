"""Which stage bounds conv v4?  Time the kernel with MMAs / A loads / B loads skipped (garbage results)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wct_tf_b200 import _capi
from tests import gpu_util as U
lib = U.lib()
B = 16
for hw, cin, cout, bn in [(512, 64, 64, 64), (256, 64, 128, 128), (64, 512, 512, 128), (64, 512, 512, 64)]:
    x = torch.randn(B, hw, hw, cin, device="cuda").relu_()
    xin = torch.empty(lib.wctb200_act_bytes(B, hw, hw, cin), dtype=torch.uint8, device="cuda")
    _capi.check(lib.wctb200_act_from_f32(x.data_ptr(), B, hw, hw, cin, xin.data_ptr(), U.stream()))
    k = (torch.randn(3, 3, cin, cout, device="cuda") * (2.0 / (9 * cin)) ** 0.5).contiguous()
    bias = torch.zeros(cout, device="cuda")
    ws = torch.empty(lib.wctb200_conv_weight_bytes(9, cin, cout), dtype=torch.uint8, device="cuda")
    _capi.check(lib.wctb200_prep_conv_weights(k.data_ptr(), 9, cin, cout, ws.data_ptr(), U.stream()))
    out = torch.empty(lib.wctb200_act_bytes(B, hw, hw, cout), dtype=torch.uint8, device="cuda")
    kiters = 9 * cin // 64 * (B * hw * hw // 128) * (cout // bn) / 148.0
    row = "%dx%d %d->%d bn%d :" % (hw, hw, cin, cout, bn)
    lib.wctb200_debug_set_conv_impl(4); lib.wctb200_debug_set_conv_bn(bn)
    for dbg in [0, 1, 2, 4, 6, 7]:
        lib.wctb200_debug_set_conv4(1, 100000 + dbg)
        def run():
            _capi.check(lib.wctb200_conv3x3(xin.data_ptr(), B, hw, hw, cin, ws.data_ptr(), bias.data_ptr(), cout, 1, out.data_ptr(), U.stream()))
        for _ in range(2): run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        row += "  dbg%d %.0fus (%.2f us/kiter)" % (dbg, ms * 1e3, ms * 1e3 / kiters)
    print(row, flush=True)
lib.wctb200_debug_set_conv4(2, 100000); lib.wctb200_debug_set_conv_impl(2); lib.wctb200_debug_set_conv_bn(0)
_capi.check(lib.wctb200_check_device(U.stream()))
