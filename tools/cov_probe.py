"""Probe the tcgen05 covariance (MN-major operands) against numpy for descriptor stride variants."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wct_tf_b200 import _capi
from tests import gpu_util as U
lib = U.lib()
rng = np.random.default_rng(0)
def run(x, impl, lbo, sbo):
    n, h, w, c = x.shape
    lib.wctb200_debug_set_cov(impl, lbo, sbo)
    buf = U.act_from_numpy(x)
    mean = torch.empty((n, c), dtype=torch.float32, device="cuda")
    cov = torch.empty((n, c, c), dtype=torch.float32, device="cuda")
    _capi.check(lib.wctb200_covariance(buf.data_ptr(), n, h, w, c, 0.0, mean.data_ptr(), cov.data_ptr(), U.stream()))
    rc = lib.wctb200_check_device(U.stream())
    return mean.cpu().numpy(), cov.cpu().numpy(), rc
for shape in [(1, 8, 32, 64), (2, 9, 40, 64), (1, 16, 33, 128), (1, 12, 20, 256), (2, 8, 8, 512)]:
    x = np.maximum(rng.standard_normal(shape) @ (rng.standard_normal((shape[3], shape[3])) / np.sqrt(shape[3])) + 0.3, 0).astype(np.float32)
    xs = U.split_repr(x).reshape(shape[0], -1, shape[3])
    ref = np.stack([np.cov(xs[i].T) for i in range(shape[0])])
    m1, c1, _ = run(x, 1, -1, -1)
    print(shape, "ffma impl: max err %.2e" % np.abs(c1 - ref).max(), flush=True)
    for lbo, sbo in [(8192, 1024), (1024, 8192), (8192, 128), (128, 1024)]:
        m2, c2, rc = run(x, 2, lbo, sbo)
        print("   tc lbo=%d sbo=%d: rc=%d max err %.2e (cov max %.2f)  mean err %.1e" % (lbo, sbo, rc, np.abs(c2 - ref).max(), np.abs(ref).max(), np.abs(m2 - xs.mean(1)).max()), flush=True)
lib.wctb200_debug_set_cov(2, 8192, 1024)
