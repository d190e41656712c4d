"""Which stage lifts null eigenvalues above 1e-5 on a rank-deficient map (HW < C)?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wct_tf_b200 import _capi
from tests import gpu_util as U
lib = U.lib()
rng = np.random.default_rng(1)
C, H, W = 512, 5, 6
x = np.maximum(rng.standard_normal((1, H, W, C)) @ (rng.standard_normal((C, C)) / np.sqrt(C)) * 1.5 + 0.8, 0).astype(np.float32)
xs = U.split_repr(x).reshape(-1, C)
cov64 = np.cov(xs.T)
w64 = np.linalg.eigvalsh(cov64)[::-1]
print("true rank", (w64 > 1e-5).sum(), "lambda_max %.2f" % w64[0])
xc32 = (x.reshape(-1, C) - x.reshape(-1, C).mean(0)).astype(np.float32)
cov32 = (xc32.T @ xc32 / np.float32(H * W - 1)).astype(np.float32)
print("numpy fp32 cov -> LAPACK fp32 svd k =", (np.linalg.svd(cov32, compute_uv=False) > 1e-5).sum(),
      "; its largest null value %.2e" % np.sort(np.linalg.svd(cov32, compute_uv=False))[::-1][29])
def gpu_cov(impl):
    lib.wctb200_debug_set_cov(impl, -1, -1)
    buf = U.act_from_numpy(x)
    mean = torch.empty((1, C), dtype=torch.float32, device="cuda"); cov = torch.empty((1, C, C), dtype=torch.float32, device="cuda")
    _capi.check(lib.wctb200_covariance(buf.data_ptr(), 1, H, W, C, 0.0, mean.data_ptr(), cov.data_ptr(), U.stream()))
    torch.cuda.synchronize(); return cov[0].cpu().numpy()
def gpu_jacobi(a):
    d = U.dev(a[None].astype(np.float32).copy()); sig = torch.empty((1, C), dtype=torch.float32, device="cuda"); sw = torch.zeros(1, dtype=torch.int32, device="cuda")
    _capi.check(lib.wctb200_jacobi_eigh(d.data_ptr(), C, 1, sig.data_ptr(), sw.data_ptr(), U.stream())); torch.cuda.synchronize()
    return np.sort(sig.cpu().numpy()[0])[::-1], int(sw.item())
for impl in (1, 2):
    c = gpu_cov(impl)
    wl = np.linalg.eigvalsh(c.astype(np.float64))[::-1]
    sg, sw = gpu_jacobi(c)
    print("gpu cov impl %d: max|cov-cov64| %.2e; exact eig of it: k=%d (null max %.2e, min %.2e); our Jacobi on it: k=%d (30th value %.2e, sweeps %d)"
          % (impl, np.abs(c - cov64).max(), (wl > 1e-5).sum(), wl[29], wl[-1], (sg > 1e-5).sum(), sg[29], sw))
sg, sw = gpu_jacobi(cov32)
print("numpy fp32 cov -> our Jacobi: k=%d (30th %.2e, 34th %.2e) sweeps %d" % ((sg > 1e-5).sum(), sg[29], sg[33], sw))
lib.wctb200_debug_set_cov(2, -1, -1)
