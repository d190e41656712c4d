"""NumPy emulation of the coupled Newton-Schulz iteration with split-fp16 x3 products: TRUE products are stable at the
rounding floor; products that use the symmetry of the operands (A.B^T) diverge after reaching ~1e-4.  (matfun_tc.cu)"""
import numpy as np
def r22(x):
    x = x.astype(np.float32); hi = x.astype(np.float16).astype(np.float32); lo = (x - hi).astype(np.float16).astype(np.float32); return hi, lo
def gemm3T(a, b):
    ah, al = r22(a); bh, bl = r22(b)
    ah, al, bh, bl = [v.astype(np.float64) for v in (ah, al, bh, bl)]
    return (ah @ bh.T + ah @ bl.T + al @ bh.T).astype(np.float32)
def gemm3(a, b):
    return gemm3T(a, b.T.copy())
def ns(A, iters, gemm, symmetrize=False):
    n = A.shape[0]; s = np.float32(np.linalg.norm(A, 'fro'))
    Y = (A / s).astype(np.float32); Z = np.eye(n, dtype=np.float32); I = np.eye(n, dtype=np.float32)
    hist = []
    for it in range(iters):
        P = gemm(Z, Y)
        hist.append(float(np.abs(I - P).max()))
        T = (1.5 * I - 0.5 * P).astype(np.float32)
        if symmetrize: T = (0.5 * (T + T.T)).astype(np.float32)
        Y = gemm(Y, T); Z = gemm(T, Z)
        if symmetrize: Y = (0.5 * (Y + Y.T)).astype(np.float32); Z = (0.5 * (Z + Z.T)).astype(np.float32)
    return hist
def cov(C, HW, seed, decay):
    r = np.random.default_rng(seed)
    X = np.maximum(r.standard_normal((C, C)) / np.sqrt(C) @ r.standard_normal((C, HW)) + 0.3, 0)
    X *= np.exp(-decay * np.arange(C) / C)[:, None]; X -= X.mean(1, keepdims=True)
    return (X @ X.T / (HW - 1) + 1e-8 * np.eye(C)).astype(np.float32)
for (C, HW, decay) in [(256, 512, 0.0), (256, 2048, 3.0)]:
    A = cov(C, HW, 1, decay); w = np.linalg.eigvalsh(A.astype(np.float64)); print("cond %.1e" % (w[-1] / w[0]))
    for name, g, sym in [("true product", gemm3, False), ("A.B^T (symmetry assumed)", gemm3T, False), ("A.B^T + symmetrise", gemm3T, True)]:
        h = ns(A, 24, g, sym)
        print("  %-28s" % name, " ".join("%.0e" % v for v in h))
