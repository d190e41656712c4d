"""Covariance stage alone (wctb200_covariance: shift sample, tcgen05 Gram with fused centring + sums, finalize): ms and
compulsory-byte GB/s per level shape, for ring depths 3 / 6 / 12.  usage: python tools/cov_bench.py [frames]"""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wct_tf_b200 import _capi
lib = _capi.load()
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 16
st = torch.cuda.current_stream().cuda_stream
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
print("%-12s" % "shape" + "".join("%26s" % ("ring<=%d" % r) for r in (3, 6, 12)) + "   (median ms, GB/s of 4*C*HW bytes, fraction of %.0f GB/s)" % peak)
for C, hw in ((64, 512), (128, 256), (256, 128), (512, 64), (512, 32)):
    feat = torch.rand((nb, hw, hw, C), dtype=torch.float32, device="cuda")
    act = torch.empty(lib.wctb200_act_bytes(nb, hw, hw, C), dtype=torch.uint8, device="cuda")
    _capi.check(lib.wctb200_act_from_f32(feat.data_ptr(), nb, hw, hw, C, act.data_ptr(), st))
    mean = torch.empty((nb, C), dtype=torch.float32, device="cuda")
    cov = torch.empty((nb, C, C), dtype=torch.float32, device="cuda")
    row = "%-12s" % ("C%d@%d" % (C, hw))
    for ring in (3, 6, 12):
        lib.wctb200_debug_set_cov_stages(ring)
        run = lambda: _capi.check(lib.wctb200_covariance(act.data_ptr(), nb, hw, hw, C, 1e-8, mean.data_ptr(), cov.data_ptr(), st))
        for _ in range(3): run()
        ts = []
        for _ in range(15):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        gbs = 4.0 * C * hw * hw * nb / (ms * 1e-3) / 1e9
        row += "%9.3f ms %6.0f (%.2f)" % (ms, gbs, gbs / peak)
    print(row, flush=True)
    del feat, act
lib.wctb200_debug_set_cov_stages(12)
_capi.check(lib.wctb200_check_device(st))
