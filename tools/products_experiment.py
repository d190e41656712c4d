"""is the third split-fp16 product needed?  For products = 3 / 2 / 1: teacher-forced per-level parity of a
5-level 256x256 frame against the fp64 oracle (the gate is 1e-3), and the conv throughput of three layer shapes."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wct_tf_b200 import _capi
from wct_tf_b200.weights import make_synthetic_weights
from tests.test_gpu_pipeline import _teacher_forced, ALL
from tests import gpu_util as U
lib = _capi.load()
w = make_synthetic_weights(42)
for n in (3, 2, 1):
    lib.wctb200_debug_set_conv_products(n)
    try:
        worst = "%.2e" % _teacher_forced(w, "tf", False, ALL, (256, 256), (256, 256), seeds=(1000, 7))
    except AssertionError as e:
        worst = "FAILED an exactness assertion (%s)" % (str(e).split("\n")[0][:60] or "k_c/k_s or level input")
    row = "products %d: worst teacher-forced error %s ;" % (n, worst)
    B = 16
    for hw, cin, cout in ((512, 64, 64), (128, 256, 256), (64, 512, 512)):
        x = torch.randn(B, hw, hw, cin, device="cuda").relu_()
        xin = torch.empty(lib.wctb200_act_bytes(B, hw, hw, cin), dtype=torch.uint8, device="cuda")
        _capi.check(lib.wctb200_act_from_f32(x.data_ptr(), B, hw, hw, cin, xin.data_ptr(), U.stream()))
        k = (torch.randn(3, 3, cin, cout, device="cuda") * (2.0 / (9 * cin)) ** 0.5).contiguous()
        bias = torch.zeros(cout, device="cuda")
        ws = torch.empty(lib.wctb200_conv_weight_bytes(9, cin, cout), dtype=torch.uint8, device="cuda")
        _capi.check(lib.wctb200_prep_conv_weights(k.data_ptr(), 9, cin, cout, ws.data_ptr(), U.stream()))
        out = torch.empty(lib.wctb200_act_bytes(B, hw, hw, cout), dtype=torch.uint8, device="cuda")
        run = lambda: _capi.check(lib.wctb200_conv3x3(xin.data_ptr(), B, hw, hw, cin, ws.data_ptr(), bias.data_ptr(), cout, 1, out.data_ptr(), U.stream()))
        for _ in range(3): run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        row += "  %dx%d %d->%d %.0f TFLOP/s" % (hw, hw, cin, cout, 2.0 * 9 * cin * cout * B * hw * hw / ms / 1e9)
    print(row, flush=True)
lib.wctb200_debug_set_conv_products(3)
