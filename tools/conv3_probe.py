"""Probe conv impl 3 (tap reuse + multicast) against the fp64 oracle for every (cluster, bo_mode)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wct_tf_b200 import _capi
from tests import gpu_util as U
from tests.test_gpu_layers import conv_ref64, _conv_inputs

CASES = [(1, 8, 8, 64, 64, True), (2, 7, 6, 128, 64, False), (1, 12, 10, 256, 256, True), (3, 34, 30, 64, 64, True),
         (1, 64, 64, 64, 128, True), (1, 8, 130, 64, 64, True), (2, 9, 128, 128, 128, True), (1, 150, 260, 64, 64, True)]
lib = U.lib()
for cluster in (1, 2):
    for bo in (1, 0):
        lib.wctb200_debug_set_conv3(cluster, bo)
        lib.wctb200_debug_set_conv_impl(3)
        for case in CASES:
            n, h, w, cin, cout, relu = case
            x, k, b = _conv_inputs(case, 7)
            xin = U.act_from_numpy(x)
            d_k, d_b = U.dev(k), U.dev(b)
            wsplit = torch.empty(lib.wctb200_conv_weight_bytes(9, cin, cout), dtype=torch.uint8, device="cuda")
            _capi.check(lib.wctb200_prep_conv_weights(d_k.data_ptr(), 9, cin, cout, wsplit.data_ptr(), U.stream()))
            out = U.act_alloc(n, h, w, cout)
            rc = lib.wctb200_conv3x3(xin.data_ptr(), n, h, w, cin, wsplit.data_ptr(), d_b.data_ptr(), cout,
                                     _capi.RELU if relu else 0, out.data_ptr(), U.stream())
            dev_rc = lib.wctb200_check_device(U.stream())
            got = U.act_to_numpy(out, n, h, w, cout)
            ref = conv_ref64(U.split_repr(x), U.split_repr(k), b, relu)
            err = np.abs(got - ref) / (1 + np.abs(ref))
            bad = int((~np.isfinite(got)).sum())
            print("cluster=%d bo=%d case=%s rc=%d dev=%d maxrel=%.2e nonfinite=%d" % (cluster, bo, case, rc, dev_rc, np.nanmax(err), bad), flush=True)
lib.wctb200_debug_set_conv3(2, 0)
lib.wctb200_debug_set_conv_impl(2)
