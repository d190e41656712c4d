"""One profiled step of the bench workload for ncu (profile-from-start off).
usage: ncu ... --profile-from-start off python tools/profile_step.py [batch]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wct_tf_b200.engine import Engine  # noqa: E402
from wct_tf_b200.weights import make_synthetic_weights  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
T = ["relu5_1", "relu4_1", "relu3_1", "relu2_1", "relu1_1"]
rng = np.random.default_rng(0)
c = torch.from_numpy(rng.integers(0, 256, (B, 512, 512, 3), dtype=np.uint8)).cuda()
s = torch.from_numpy(rng.integers(0, 256, (B, 512, 512, 3), dtype=np.uint8)).cuda()
eng = Engine(make_synthetic_weights(42), T, semantics="tf")
eng.to_u8(eng.stylize(c, s, alpha=0.8))
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
eng.to_u8(eng.stylize(c, s, alpha=0.8))
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
eng.check_device()
print("profiled one step, batch", B)
