"""k_c, k_s (eigenvalues kept above the 1e-5 threshold, ops.py:57-64) and Jacobi sweeps per level for the bench workload
(uniform-noise frames, seeded synthetic weights) and for a smooth natural-ish input."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from wct_tf_b200.engine import Engine
from wct_tf_b200.weights import make_synthetic_weights
targets = ["relu5_1", "relu4_1", "relu3_1", "relu2_1", "relu1_1"]
eng = Engine(make_synthetic_weights(42, relu_targets=targets), targets)
rng = np.random.default_rng(0)
def smooth(seed):
    r = np.random.default_rng(seed).standard_normal((66, 66, 3))
    big = np.kron(r, np.ones((8, 8, 1)))[:512, :512]
    return np.uint8(np.clip(128 + 60 * big, 0, 255))
for name, c, s in [("noise", rng.integers(0, 256, (2, 512, 512, 3), dtype=np.uint8), rng.integers(0, 256, (2, 512, 512, 3), dtype=np.uint8)),
                   ("smooth", np.stack([smooth(1), smooth(2)]), np.stack([smooth(3), smooth(4)]))]:
    eng.stylize(torch.from_numpy(c).cuda(), torch.from_numpy(s).cuda(), alpha=0.8, want_info=True)
    print(name)
    for info in eng.last_info:
        print("   ", info)
