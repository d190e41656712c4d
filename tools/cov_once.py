"""One wctb200_covariance call per level shape between cudaProfilerStart/Stop (for an ncu launch list)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wct_tf_b200 import _capi
lib = _capi.load()
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 16
st = torch.cuda.current_stream().cuda_stream
runs = []
for C, hw in ((64, 512), (128, 256), (256, 128), (512, 64)):
    feat = torch.rand((nb, hw, hw, C), dtype=torch.float32, device="cuda")
    act = torch.empty(lib.wctb200_act_bytes(nb, hw, hw, C), dtype=torch.uint8, device="cuda")
    _capi.check(lib.wctb200_act_from_f32(feat.data_ptr(), nb, hw, hw, C, act.data_ptr(), st))
    mean = torch.empty((nb, C), dtype=torch.float32, device="cuda")
    cov = torch.empty((nb, C, C), dtype=torch.float32, device="cuda")
    runs.append((act, mean, cov, C, hw))
def go():
    for act, mean, cov, C, hw in runs:
        _capi.check(lib.wctb200_covariance(act.data_ptr(), nb, hw, hw, C, 1e-8, mean.data_ptr(), cov.data_ptr(), st))
go(); go(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
go(); torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
