"""Matrix-function fast path (matfun_tc.cu): accuracy against float64 eigh and time against the Jacobi path.
python tools/matfun_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from wct_tf_b200 import _capi
from tests import gpu_util as U

lib = _capi.load()


def covs(C, HW, count, decay, seed, dead=0):
    out = []
    for i in range(count):
        r = np.random.default_rng(seed + i)
        M = r.standard_normal((C, C)) / np.sqrt(C)
        X = np.maximum(M @ r.standard_normal((C, HW)) + 0.3, 0)
        X *= np.exp(-decay * np.arange(C) / C)[:, None]
        if dead:
            X[r.choice(C, dead, replace=False)] = 0.0
        X -= X.mean(1, keepdims=True)
        out.append((X @ X.T / (HW - 1) + 1e-8 * np.eye(C)).astype(np.float32))
    return np.stack(out)


def mfun(A, p):
    w, v = np.linalg.eigh(A.astype(np.float64))
    return (v * w ** p) @ v.T, w


def timeit(f, n=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for C, HW, count in [(512, 1024, 15), (512, 4096, 15), (256, 4096, 15), (128, 16384, 15)]:
    for decay, dead in [(0.0, 0), (3.0, 0), (4.5, 0), (0.0, 8)]:
        A = covs(C, HW, count, decay, 10 * C + int(decay * 10), dead)
        dA = U.dev(A)
        out = torch.zeros_like(dA)
        ok = torch.zeros(count, dtype=torch.int32, device="cuda")
        nf = count // 2
        info = torch.zeros(count * 4, dtype=torch.float32, device="cuda")
        f = lambda: _capi.check(lib.wctb200_debug_matfun(dA.data_ptr(), C, count, nf, 1e-5, 0.0, out.data_ptr(), ok.data_ptr(), info.data_ptr(), U.stream()))
        ms = timeit(f)
        okh = ok.cpu().numpy()
        o = out.cpu().numpy()
        errs, conds = [], []
        for b in range(count):
            ref, w = mfun(A[b], -0.5 if b < nf else 0.5)
            conds.append(w[-1] / max(w[0], 1e-30))
            if okh[b]:
                errs.append(np.abs(o[b] - ref).max() / np.abs(ref).max())
        print("C=%3d HW=%5d decay %.1f dead %d: cond %.1e..%.1e  lambda_min %.1e  ok %2d/%d  max rel err %s  %.3f ms / %d matrices"
              % (C, HW, decay, dead, min(conds), max(conds), min(np.linalg.eigvalsh(A[0].astype(np.float64))), okh.sum(), count,
                 ("%.1e" % max(errs)) if errs else "-", ms, count), flush=True)
        ih = info.cpu().numpy().reshape(count, 4)
        print("      converged at iteration %s, last residual %.1e..%.1e, lambda_min bound %.1e..%.1e" % (sorted(set(int(v) for v in ih[:, 0])), ih[:, 1].min(), ih[:, 1].max(), ih[:, 2].min(), ih[:, 2].max()), flush=True)
# Jacobi for comparison
for C in (512, 256):
    A = covs(C, 1024 if C == 512 else 4096, 15, 0.0, 1)
    d = U.dev(A); sw = torch.zeros(15, dtype=torch.int32, device="cuda"); sg = torch.zeros(15 * C, dtype=torch.float32, device="cuda")
    def g():
        t = d.clone()
        _capi.check(lib.wctb200_jacobi_eigh(t.data_ptr(), C, 15, sg.data_ptr(), sw.data_ptr(), U.stream()))
    print("Jacobi C=%d x15 (incl. a device copy of the input): %.3f ms" % (C, timeit(g)))
