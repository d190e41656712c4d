"""Timeline of CTA 0 of conv v4 (clock64 samples): where does a k-iteration spend its time?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wct_tf_b200 import _capi
from tests import gpu_util as U
lib = U.lib()
B = 16
for hw, cin, cout, bn, dbg in [(64, 512, 512, 128, 0), (64, 512, 512, 64, 0), (64, 512, 512, 64, 7), (512, 64, 64, 64, 0)]:
    x = torch.randn(B, hw, hw, cin, device="cuda").relu_()
    xin = torch.empty(lib.wctb200_act_bytes(B, hw, hw, cin), dtype=torch.uint8, device="cuda")
    _capi.check(lib.wctb200_act_from_f32(x.data_ptr(), B, hw, hw, cin, xin.data_ptr(), U.stream()))
    k = (torch.randn(3, 3, cin, cout, device="cuda") * (2.0 / (9 * cin)) ** 0.5).contiguous()
    bias = torch.zeros(cout, device="cuda")
    ws = torch.empty(lib.wctb200_conv_weight_bytes(9, cin, cout), dtype=torch.uint8, device="cuda")
    _capi.check(lib.wctb200_prep_conv_weights(k.data_ptr(), 9, cin, cout, ws.data_ptr(), U.stream()))
    out = torch.empty(lib.wctb200_act_bytes(B, hw, hw, cout), dtype=torch.uint8, device="cuda")
    lib.wctb200_debug_set_conv_impl(4); lib.wctb200_debug_set_conv_bn(bn); lib.wctb200_debug_set_conv4(1, 100000 + dbg)
    tr = torch.zeros(1024, dtype=torch.int64, device="cuda")
    for rep in range(2):
        lib.wctb200_debug_conv4_trace(tr.data_ptr() if rep == 1 else None)
        _capi.check(lib.wctb200_conv3x3(xin.data_ptr(), B, hw, hw, cin, ws.data_ptr(), bias.data_ptr(), cout, 1, out.data_ptr(), U.stream()))
        torch.cuda.synchronize()
    lib.wctb200_debug_conv4_trace(None)
    t = tr.cpu().numpy()
    st, rdy, done = t[0:256], t[256:512], t[512:768]
    ch = t[768:1024].reshape(128, 2)
    t0 = st[0]
    print("=== %dx%d %d->%d bn%d dbg%d  (cycles; k-iter: start->operands ready->issued ; period)" % (hw, hw, cin, cout, bn, dbg))
    for i in list(range(0, 14)) + list(range(100, 112)):
        print("  kiter %3d  t=%8d  wait %6d  issue %5d  period %6d" % (i, st[i] - t0, rdy[i] - st[i], done[i] - rdy[i], (st[i + 1] - st[i]) if i < 255 else 0))
    per = np.diff(st[40:250])
    print("  steady k-iter period: mean %.0f median %.0f ; wait mean %.0f ; issue mean %.0f" % (per.mean(), np.median(per), (rdy - st)[40:250].mean(), (done - rdy)[40:250].mean()))
    cs, cd = ch[:, 0], ch[:, 1]
    print("  chunks: seen->drained mean %.0f ; chunk period mean %.0f ; first chunks (seen-t0, drain):" % ((cd - cs)[10:60].mean(), np.diff(cs[10:60]).mean()),
          [(int(cs[i] - t0), int(cd[i] - cs[i])) for i in range(6)])
lib.wctb200_debug_set_conv4(2, 100000); lib.wctb200_debug_set_conv_impl(2); lib.wctb200_debug_set_conv_bn(0)
_capi.check(lib.wctb200_check_device(U.stream()))
