"""Serial (overlap off) per-call breakdown of one bench step via the engine's CUDA-event profile hooks.
usage: python tools/step_breakdown.py [batch]"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wct_tf_b200.engine import Engine
from wct_tf_b200.weights import make_synthetic_weights

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = ["relu5_1", "relu4_1", "relu3_1", "relu2_1", "relu1_1"]
rng = np.random.default_rng(0)
c = torch.from_numpy(rng.integers(0, 256, (B, 512, 512, 3), dtype=np.uint8)).cuda()
s = torch.from_numpy(rng.integers(0, 256, (B, 512, 512, 3), dtype=np.uint8)).cuda()
eng = Engine(make_synthetic_weights(42), T, semantics="tf")
eng.overlap_style = False
eng.groups = 1
for _ in range(2):
    eng.to_u8(eng.stylize(c, s, alpha=0.8))
torch.cuda.synchronize()
eng.profile = {}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
eng.to_u8(eng.stylize(c, s, alpha=0.8))
e1.record()
torch.cuda.synchronize()
tot = e0.elapsed_time(e1)
rows = []
for k, rec in eng.profile.items():
    ms = sum(a.elapsed_time(b) for a, b in rec["events"])
    rows.append((ms, k, len(rec["events"]), rec["flops"], rec["bytes"]))
rows.sort(reverse=True)
acc = sum(r[0] for r in rows)
print("batch %d serial step %.2f ms ; sum of calls %.2f ms" % (B, tot, acc))
for ms, k, n, fl, by in rows:
    extra = ""
    if fl: extra = " %.0f TFLOP/s" % (fl / ms / 1e9)
    elif by: extra = " %.0f GB/s" % (by / ms / 1e6)
    print("%-28s n=%3d %8.3f ms %5.1f%%%s" % (k, n, ms, 100 * ms / acc, extra))

# ---- overlapped schedule: host enqueue time vs device time
import time
eng.profile = None
eng.overlap_style = True
eng.groups = 4
for _ in range(2):
    eng.to_u8(eng.stylize(c, s, alpha=0.8))
torch.cuda.synchronize()
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    out = eng.to_u8(eng.stylize(c, s, alpha=0.8))
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("overlapped step: host enqueue %.1f ms, host until done %.1f ms, device %.1f ms, launches %d" % (
        (t1 - t0) * 1e3, (t2 - t0) * 1e3, e0.elapsed_time(e1), getattr(eng, "launches", -1)))
