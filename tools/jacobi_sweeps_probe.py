"""CPU probe: sweeps of a cyclic one-sided Jacobi (round-robin ordering, the kernel's stopping rule) on a covariance of
relu-like features, plain vs column-sorted vs pivoted-Cholesky-preconditioned.  python tools/jacobi_sweeps_probe.py C HW"""
import numpy as np, sys
rng = np.random.default_rng(0)
def feats(C, HW, seed):
    r = np.random.default_rng(seed)
    M = r.standard_normal((C, C)).astype(np.float32) / np.sqrt(C)
    X = np.maximum(M @ r.standard_normal((C, HW)).astype(np.float32) + 0.3, 0)
    return X
def cov(X):
    Xc = X - X.mean(1, keepdims=True)
    return (Xc @ Xc.T / (X.shape[1] - 1) + 1e-8 * np.eye(X.shape[0])).astype(np.float32)
def rr_rounds(n):
    idx = list(range(n))
    for r in range(n - 1):
        yield [(idx[i], idx[n - 1 - i]) for i in range(n // 2)]
        idx = [idx[0]] + [idx[-1]] + idx[1:-1]
def jacobi(G, tol, maxs=40):
    G = G.astype(np.float32).copy(); n = G.shape[1]
    for sweep in range(1, maxs + 1):
        worst = 0.0; nrot = 0
        for pairs in rr_rounds(n):
            p = np.array(pairs); i, j = p[:, 0], p[:, 1]
            x, y = G[:, i], G[:, j]
            a = (x * x).sum(0); b = (y * y).sum(0); c = (x * y).sum(0)
            rel = np.abs(c) / np.sqrt(a * b + 1e-300)
            worst = max(worst, rel.max())
            act = rel > tol
            nrot += act.sum()
            zeta = (b - a) / (2 * np.where(c == 0, 1, c))
            t = np.sign(zeta) / (np.abs(zeta) + np.sqrt(1 + zeta * zeta)); t = np.where(zeta == 0, 1.0, t)
            t = np.where(act, t, 0)
            cs = 1 / np.sqrt(1 + t * t); sn = cs * t
            G[:, i] = (cs * x - sn * y).astype(np.float32); G[:, j] = (sn * x + cs * y).astype(np.float32)
        print("   sweep %d worst %.2e rotations %d" % (sweep, worst, nrot), flush=True)
        if worst <= tol: return sweep, G
    return maxs, G
def pivchol(A):
    A = A.astype(np.float64).copy(); n = A.shape[0]; L = np.zeros((n, n)); perm = np.arange(n)
    d = np.diag(A).copy()
    for k in range(n):
        p = k + np.argmax(d[k:])
        if d[p] <= 0: break
        perm[[k, p]] = perm[[p, k]]; L[[k, p], :] = L[[p, k], :]; d[[k, p]] = d[[p, k]]
        L[k, k] = np.sqrt(d[k])
        col = (A[perm[k+1:], perm[k]] - L[k+1:, :k] @ L[k, :k]) / L[k, k]
        L[k+1:, k] = col; d[k+1:] -= col * col
    return L.astype(np.float32), perm
C = int(sys.argv[1]); HW = int(sys.argv[2]); tol = 2 * np.sqrt(C) * 5.96e-8
A = cov(feats(C, HW, 3))
w = np.linalg.eigvalsh(A.astype(np.float64)); print("eig range", w[0], w[-1], "k>1e-5", (w > 1e-5).sum())
print("plain G=A"); s, G = jacobi(A, tol)
nr = np.sort((G * G).sum(0) ** 0.5)[::-1]; print("  sweeps", s, "eig err", np.abs(nr - w[::-1]).max() / w[-1])
print("sorted columns"); o = np.argsort(-(A * A).sum(0)); s, G = jacobi(A[:, o], tol); print("  sweeps", s)
print("pivoted cholesky factor L (columns of L)"); L, perm = pivchol(A); s, G = jacobi(L, tol)
nr = np.sort((G * G).sum(0))[::-1]; print("  sweeps", s, "eig err", np.abs(nr - w[::-1]).max() / w[-1])
print("L^T (rows of L)"); s, G = jacobi(L.T.copy(), tol); print("  sweeps", s)
