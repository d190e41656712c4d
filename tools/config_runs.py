"""BASELINE.json configs 4 and 5 (and the batch-64 shard of config 3) on one GPU: timing only."""
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wct_tf_b200.engine import Engine
from wct_tf_b200.weights import make_synthetic_weights

T = ["relu5_1", "relu4_1", "relu3_1", "relu2_1", "relu1_1"]
eng = Engine(make_synthetic_weights(42), T, semantics="tf")
rng = np.random.default_rng(0)
def u8(n, s): return torch.from_numpy(rng.integers(0, 256, (n, s, s, 3), dtype=np.uint8)).cuda()
def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
out = {}
c1, s1 = u8(1, 512), u8(1, 512)
out["config2_single_frame_512_ms"] = timeit(lambda: eng.to_u8(eng.stylize(c1, s1, alpha=0.8)))
c4, s4 = u8(1, 1024), u8(1, 512)
out["config4_1024_content_512_style_ms"] = timeit(lambda: eng.to_u8(eng.stylize(c4, s4, alpha=0.8)))
c5, s5 = u8(8, 512), u8(8, 512)
out["config5_adain_batch8_ms"] = timeit(lambda: eng.to_u8(eng.stylize(c5, s5, alpha=0.8, adain=True)))
out["config5_adain_fps"] = 8000.0 / out["config5_adain_batch8_ms"]
c3, s3 = u8(8, 512), u8(1, 512)
out["config3_shard_8_frames_shared_style_ms"] = timeit(lambda: eng.to_u8(eng.stylize(c3, s3, alpha=0.8)))
out["config3_shared_style_fps_per_gpu"] = 8000.0 / out["config3_shard_8_frames_shared_style_ms"]
eng.check_device()
print(json.dumps(out, indent=1))
