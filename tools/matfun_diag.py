import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from wct_tf_b200 import _capi
from tests import gpu_util as U
lib = _capi.load()
for C in (128, 256, 512):
    r = np.random.default_rng(C)
    X = np.maximum(r.standard_normal((C, C)) / np.sqrt(C) @ r.standard_normal((C, 2048)) + 0.3, 0); X -= X.mean(1, keepdims=True)
    A = np.stack([(X @ X.T / 2047).astype(np.float32)] * 4)
    dA = U.dev(A); out = torch.zeros_like(dA); ok = torch.zeros(4, dtype=torch.int32, device="cuda")
    for iters in (1, 2):
        lib.wctb200_debug_set_matfun(1, iters)
        _capi.check(lib.wctb200_debug_matfun(dA.data_ptr(), C, 4, 2, 1e-5, 0.0, out.data_ptr(), ok.data_ptr(), None, U.stream()))
        o = out.cpu().numpy().astype(np.float64)
        a = A[0].astype(np.float64); s = np.linalg.norm(a); I = np.eye(C)
        Y, Z = a / s, I
        for _ in range(iters):
            T = 1.5 * I - 0.5 * Z @ Y
            Y, Z = Y @ T, T @ Z
        ez = np.abs(o[0] - Z / np.sqrt(s)); ey = np.abs(o[3] - Y * np.sqrt(s))
        print("C=%d iters %d: Z err max %.2e (ref max %.2e)  Y err max %.2e (ref max %.2e)" % (C, iters, ez.max(), np.abs(Z / np.sqrt(s)).max(), ey.max(), np.abs(Y * np.sqrt(s)).max()))
        if ez.max() > 1e-3:
            bad = np.argwhere(ez > 1e-3)
            print("    bad Z entries: rows %d..%d cols %d..%d count %d" % (bad[:, 0].min(), bad[:, 0].max(), bad[:, 1].min(), bad[:, 1].max(), len(bad)))
            tiles = sorted(set((int(i) // 128, int(j) // 128) for i, j in bad))
            print("    bad tiles", tiles[:20])
lib.wctb200_debug_set_matfun(1, 16)
