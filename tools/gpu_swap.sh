#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_transform.py -q -m gpu -x -s -k "style_swap" 2>&1 | tail -6
timeout 300 python - <<'PY' 2>&1 | tail -6
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from wct_tf_b200.wct import WCT
from wct_tf_b200.weights import make_synthetic_weights
T = ["relu5_1", "relu4_1", "relu3_1", "relu2_1", "relu1_1"]
w = make_synthetic_weights(42)
wct = WCT(checkpoints=None, relu_targets=T, vgg_path=None, weights=w)
rng = np.random.default_rng(0)
c = rng.integers(0, 256, (512, 512, 3), dtype=np.uint8); s = rng.integers(0, 256, (512, 512, 3), dtype=np.uint8)
for _ in range(2): out = wct.predict(c, s, alpha=0.8, swap5=True, ss_alpha=0.6)
torch.cuda.synchronize(); t0 = time.time()
for _ in range(5): out = wct.predict(c, s, alpha=0.8, swap5=True, ss_alpha=0.6)
torch.cuda.synchronize(); print("swap5 5-level 512x512 single frame: %.1f ms (end to end, host buffers)" % ((time.time() - t0) / 5 * 1e3), out.shape, out.dtype)
for _ in range(2): out = wct.predict(c, s, alpha=0.8)
torch.cuda.synchronize(); t0 = time.time()
for _ in range(5): out = wct.predict(c, s, alpha=0.8)
torch.cuda.synchronize(); print("plain WCT same frame: %.1f ms" % ((time.time() - t0) / 5 * 1e3))
wct.engine.check_device()
PY
