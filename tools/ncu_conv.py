"""One conv launch per shape for `ncu --set full` (profile-from-start off): 3x3 64->64 @512, UP2 64->64 @512, 3x3 512->512 @64,
3x3 256->256 @128 (the layer shape with the largest share of the step).  usage: python tools/ncu_conv.py [frames]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wct_tf_b200 import _capi
from tests import gpu_util as U
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
lib = U.lib()
def mk(hw, cin, cout, up2):
    hin = hw // 2 if up2 else hw
    x = torch.randn(B, hin, hin, cin, device="cuda").relu_()
    xin = torch.empty(lib.wctb200_act_bytes(B, hin, hin, cin), dtype=torch.uint8, device="cuda")
    _capi.check(lib.wctb200_act_from_f32(x.data_ptr(), B, hin, hin, cin, xin.data_ptr(), U.stream()))
    k = (torch.randn(3, 3, cin, cout, device="cuda") * (2.0 / (9 * cin)) ** 0.5).contiguous()
    bias = torch.zeros(cout, device="cuda")
    ws = torch.empty(lib.wctb200_conv_weight_bytes(16 if up2 else 9, cin, cout), dtype=torch.uint8, device="cuda")
    if up2: _capi.check(lib.wctb200_prep_conv_weights_up2(k.data_ptr(), cin, cout, ws.data_ptr(), U.stream()))
    else: _capi.check(lib.wctb200_prep_conv_weights(k.data_ptr(), 9, cin, cout, ws.data_ptr(), U.stream()))
    out = torch.empty(lib.wctb200_act_bytes(B, hw, hw, cout), dtype=torch.uint8, device="cuda")
    fn = lib.wctb200_conv3x3_up2 if up2 else lib.wctb200_conv3x3
    return lambda: _capi.check(fn(xin.data_ptr(), B, hin, hin, cin, ws.data_ptr(), bias.data_ptr(), cout, 1, out.data_ptr(), U.stream())), (xin, ws, bias, out, k)
runs = [mk(512, 64, 64, 0), mk(512, 64, 64, 1), mk(64, 512, 512, 0), mk(128, 256, 256, 0)]
for r, _ in runs: r()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for r, _ in runs: r()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
