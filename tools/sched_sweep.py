"""frames/s of the resident 5-level step for (frames per step, groups, conv grid over-subscription, priorities).
python tools/sched_sweep.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from wct_tf_b200.engine import Engine
from wct_tf_b200.weights import make_synthetic_weights

targets = ["relu5_1", "relu4_1", "relu3_1", "relu2_1", "relu1_1"]
eng = Engine(make_synthetic_weights(42, relu_targets=targets), targets)
rng = np.random.default_rng(0)
cs = {}
def run(B, G, oversub, prio, steps=4):
    if B not in cs:
        cs[B] = (torch.from_numpy(rng.integers(0, 256, (B, 512, 512, 3), dtype=np.uint8)).cuda(),
                 torch.from_numpy(rng.integers(0, 256, (B, 512, 512, 3), dtype=np.uint8)).cuda())
    c, s = cs[B]
    eng.groups, eng.group_priorities = G, bool(prio)
    eng.lib.wctb200_debug_set_conv_oversub(oversub)
    for _ in range(2):
        eng.stylize(c, s, alpha=0.8)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.stylize(c, s, alpha=0.8)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print("frames %3d groups %d oversub %2d prio %d : %7.2f ms/step  %6.1f frames/s" % (B, G, oversub, prio, ms, B / ms * 1e3), flush=True)

for (B, G) in [(30, 2), (45, 3), (60, 4), (90, 6), (15, 1), (30, 1)]:
    run(B, G, 4, 1)
for ov in (2, 8, 16):
    run(30, 2, ov, 1)
run(60, 4, 8, 1)
run(60, 4, 4, 0)
eng.lib.wctb200_debug_set_conv_oversub(4)
