"""One wctb200_jacobi_eigh launch (C = 512, 4 feature-like covariances) between cudaProfilerStart/Stop, for ncu."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wct_tf_b200 import _capi
lib = _capi.load()
st = torch.cuda.current_stream().cuda_stream
C, hw, count = 512, 1024, 4
rng = np.random.default_rng(C + hw)
mats = []
for i in range(count):
    m = rng.standard_normal((C, C)) / np.sqrt(C)
    x = np.maximum(rng.standard_normal((hw, C)) @ m + 0.3, 0.0)
    x = x - x.mean(0)
    mats.append((x.T @ x / (hw - 1)).astype(np.float32))
a0 = torch.from_numpy(np.stack(mats)).cuda()
sigma = torch.empty(count, C, device="cuda")
sweeps = torch.zeros(count, dtype=torch.int32, device="cuda")
a = a0.clone()
_capi.check(lib.wctb200_jacobi_eigh(a.data_ptr(), C, count, sigma.data_ptr(), sweeps.data_ptr(), st))
torch.cuda.synchronize()
a = a0.clone()
torch.cuda.cudart().cudaProfilerStart()
_capi.check(lib.wctb200_jacobi_eigh(a.data_ptr(), C, count, sigma.data_ptr(), sweeps.data_ptr(), st))
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("sweeps", sweeps.cpu().tolist())
