"""Time wctb200_jacobi_eigh alone: feature-like covariances (relu(N(0,1) M + 0.3)), per-launch ms and sweeps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from wct_tf_b200 import _capi

lib = _capi.load()
st = torch.cuda.current_stream().cuda_stream
rng = np.random.default_rng(0)
SCHED = [(int(a), int(b)) for a, b in (x.split(":") for x in os.environ.get("JAC_SCHED", "-1:-1").split(","))]
CASES = [(64, 4096, 16), (128, 4096, 16), (256, 2048, 16), (512, 1024, 15), (512, 1024, 4), (512, 300, 4)]
if os.environ.get("JAC_CASES"):
    CASES = [tuple(int(v) for v in c.split(":")) for c in os.environ["JAC_CASES"].split(",")]
for (C, hw, count), (lg, stag) in [(c, s) for c in CASES for s in SCHED]:
    lib.wctb200_debug_set_jacobi(lg, stag)
    if os.environ.get("JAC_TOLQ"):
        lib.wctb200_debug_set_jacobi_tolq(float(os.environ["JAC_TOLQ"]))
    mats = []
    rng = np.random.default_rng(C + hw)
    for i in range(count):
        m = rng.standard_normal((C, C)) / np.sqrt(C)
        x = np.maximum(rng.standard_normal((hw, C)) @ m + 0.3, 0.0)
        x = x - x.mean(0)
        mats.append((x.T @ x / (hw - 1)).astype(np.float32))
    a0 = torch.from_numpy(np.stack(mats)).cuda()
    sigma = torch.empty(count, C, device="cuda")
    sweeps = torch.zeros(count, dtype=torch.int32, device="cuda")
    ts = []
    for it in range(4):
        a = a0.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        _capi.check(lib.wctb200_jacobi_eigh(a.data_ptr(), C, count, sigma.data_ptr(), sweeps.data_ptr(), st))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    w = np.linalg.eigvalsh(np.stack(mats).astype(np.float64))[:, ::-1]
    got = np.sort(sigma.cpu().numpy(), axis=1)[:, ::-1]
    err = np.abs(got - np.abs(w)).max() / w.max()
    g = a.cpu().numpy().astype(np.float64)      # columns sigma_i u_i : check orthogonality of the normalised columns
    worst = 0.0
    for i in range(min(count, 2)):
        q = g[i] / np.maximum(np.linalg.norm(g[i], axis=1, keepdims=True), 1e-300)
        big = np.linalg.norm(g[i], axis=1) > 1e-3 * w.max()
        qq = q[big] @ q[big].T
        worst = max(worst, np.abs(qq - np.eye(qq.shape[0])).max())
    print("lg=%d stagger=%d C=%d hw=%d count=%d  ms/launch %s  sweeps %s  eig err/lmax %.2e  orth %.2e" % (
        lg, stag, C, hw, count, ["%.2f" % t for t in ts[1:]], sorted(set(sweeps.cpu().tolist())), err, worst), flush=True)
