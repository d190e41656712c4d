"""Timeline of ONE overlapped bench step (groups of sub-batches on stream pairs): every C-ABI call bracketed by CUDA events
on its own stream, nothing serialised.  Prints per stream the calls with start/end (ms from the step start), and a coarse
concurrency profile.  python tools/timeline.py [frames] [groups]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from wct_tf_b200.engine import Engine
from wct_tf_b200.weights import make_synthetic_weights

B = int(sys.argv[1]) if len(sys.argv) > 1 else 30
G = int(sys.argv[2]) if len(sys.argv) > 2 else 2
PRIO = int(sys.argv[3]) if len(sys.argv) > 3 else 1
targets = ["relu5_1", "relu4_1", "relu3_1", "relu2_1", "relu1_1"]
eng = Engine(make_synthetic_weights(42, relu_targets=targets), targets)
eng.groups, eng.group_priorities = G, bool(PRIO)
rng = np.random.default_rng(0)
c = torch.from_numpy(rng.integers(0, 256, (B, 512, 512, 3), dtype=np.uint8)).cuda()
s = torch.from_numpy(rng.integers(0, 256, (B, 512, 512, 3), dtype=np.uint8)).cuda()
for _ in range(3):
    eng.stylize(c, s, alpha=0.8)
torch.cuda.synchronize()

calls = []
orig = eng._call
def traced(key, nk, fn, *args, flops=0.0, bytes_=0.0):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.current_stream(eng.device)
    e0.record(st)
    orig(key, nk, fn, *args, flops=flops, bytes_=bytes_)
    e1.record(st)
    calls.append((eng._group, st.cuda_stream, key, e0, e1))
eng._call = traced
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
eng.stylize(c, s, alpha=0.8)
t1.record()
torch.cuda.synchronize()
total = t0.elapsed_time(t1)
rows = [(g, st, k, t0.elapsed_time(a), t0.elapsed_time(b)) for g, st, k, a, b in calls]
print("step %.2f ms, %d calls, %d frames, %d groups, priorities %d" % (total, len(rows), B, G, PRIO))
streams = sorted(set(r[1] for r in rows))
for st in streams:
    rs = [r for r in rows if r[1] == st]
    busy = sum(r[4] - r[3] for r in rs)
    print("stream %x group %d: %d calls, first %.2f last %.2f, sum of call spans %.2f ms" % (st, rs[0][0], len(rs), rs[0][3], rs[-1][4], busy))
print("--- big calls (>= 2 ms span)")
for r in sorted(rows, key=lambda r: r[3]):
    if r[4] - r[3] >= 2.0:
        print("  g%d %-28s %8.2f -> %8.2f  (%6.2f ms)" % (r[0], r[2], r[3], r[4], r[4] - r[3]))
# concurrency: at 0.25 ms resolution, how many streams have a call in flight; and whether a C512 transform is in flight
res = 0.25
nb = int(total / res) + 1
act = np.zeros(nb, int); jac = np.zeros(nb, int)
for r in rows:
    a, b = int(r[3] / res), int(r[4] / res)
    act[a:b + 1] += 1
    if "C512" in r[2]:
        jac[a:b + 1] += 1
print("--- time with k calls in flight (ms):", {int(k): round(float((act == k).sum() * res), 1) for k in np.unique(act)})
print("--- time with k C512 transforms in flight (ms):", {int(k): round(float((jac == k).sum() * res), 1) for k in np.unique(jac)})
print("--- time with a C512 transform in flight and NO other call:", round(float(((jac >= 1) & (act == jac)).sum() * res), 1), "ms")
# idle gaps: time during which NO call is in flight on any stream, and the largest per-stream gaps
ev = sorted([(r[3], 1) for r in rows] + [(r[4], -1) for r in rows])
idle, depth, last = 0.0, 0, 0.0
for t, d in ev:
    if depth == 0:
        idle += t - last
    depth += d
    last = t
print("--- time with NO call in flight: %.2f ms of %.2f" % (idle, total))
for st in streams:
    rs = sorted([r for r in rows if r[1] == st], key=lambda r: r[3])
    gaps = [(rs[i + 1][3] - rs[i][4], rs[i][2], rs[i + 1][2], rs[i][4]) for i in range(len(rs) - 1)]
    gaps.sort(reverse=True)
    print("stream %x: sum of gaps between consecutive calls %.2f ms; largest:" % (st, sum(g[0] for g in gaps)))
    for g in gaps[:6]:
        print("     %.2f ms at t=%.2f after %s before %s" % (g[0], g[3], g[1], g[2]))
if "--all" in sys.argv:
    for r in sorted(rows, key=lambda r: r[3]):
        print("  g%d %x %-30s %8.2f -> %8.2f  (%6.3f ms)" % (r[0], r[1] & 0xfff, r[2], r[3], r[4], r[4] - r[3]))
