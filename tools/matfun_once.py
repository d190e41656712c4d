"""one call of the matrix-function fast path on 15 well-conditioned 512x512 covariances (for ncu)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from wct_tf_b200 import _capi
from tests import gpu_util as U
lib = _capi.load()
C, count = 512, 15
mats = []
for i in range(count):
    r = np.random.default_rng(i)
    X = np.maximum(r.standard_normal((C, C)) / np.sqrt(C) @ r.standard_normal((C, 2048)) + 0.3, 0); X -= X.mean(1, keepdims=True)
    mats.append((X @ X.T / 2047 + 1e-8 * np.eye(C)).astype(np.float32))
dA = U.dev(np.stack(mats)); out = torch.zeros_like(dA); ok = torch.zeros(count, dtype=torch.int32, device="cuda")
for _ in range(2):
    _capi.check(lib.wctb200_debug_matfun(dA.data_ptr(), C, count, 8, 1e-5, 0.0, out.data_ptr(), ok.data_ptr(), None, U.stream()))
torch.cuda.synchronize()
print("ok", ok.cpu().tolist())
