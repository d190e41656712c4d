"""a few launches of the tensor-core tail and head kernels (for ncu).  python tools/tail_head_once.py [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from wct_tf_b200 import _capi
from tests import gpu_util as U
lib = _capi.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
h = w = 512
rng = np.random.default_rng(0)
img = torch.from_numpy(rng.random((n, h, w, 3)).astype(np.float32)).cuda()
kh = torch.from_numpy((rng.normal(0, 1, (27, 64)) * 0.2).astype(np.float32)).cuda()
bh = torch.from_numpy(rng.normal(0, 0.1, 64).astype(np.float32)).cuda()
act = torch.empty(lib.wctb200_act_bytes(n, h, w, 64), dtype=torch.uint8, device="cuda")
kt = torch.from_numpy((rng.normal(0, 1, (576, 3)) * 0.05).astype(np.float32)).cuda()
bt = torch.tensor([0.5, 0.4, 0.6], device="cuda")
out = torch.empty((n, h, w, 3), dtype=torch.float32, device="cuda")
for _ in range(2):
    _capi.check(lib.wctb200_conv_head(img.data_ptr(), n, h, w, kh.data_ptr(), bh.data_ptr(), act.data_ptr(), U.stream()))
    _capi.check(lib.wctb200_conv_tail(act.data_ptr(), n, h, w, 64, kt.data_ptr(), bt.data_ptr(), _capi.CLIP01, out.data_ptr(), U.stream()))
torch.cuda.synchronize()
print("ok", float(out.mean()))
