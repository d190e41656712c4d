"""Frame sharding across the GPUs of one box.

Frames are independent (``predict`` is stateless between calls, wct.py:97-103), so a
batch shards with NO collective on the hot path: frame i of B goes to rank floor(i*G/B)
(contiguous blocks); weights are replicated at init.  One ``all_gather`` of the finished
uint8 frames (786 KB each at 512x512) returns the batch to every rank -- NCCL over
NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_frames, world, rank):
    """Contiguous block of frames owned by ``rank``: frame i -> rank floor(i*world/n_frames)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank %r/%r" % (world, rank))
    lo = (rank * n_frames + world - 1) // world
    hi = ((rank + 1) * n_frames + world - 1) // world
    return lo, hi


def owner_of(i, n_frames, world):
    return (i * world) // n_frames


def stylize_sharded(process_fn, contents, styles, group=None):
    """Run ``process_fn(contents_shard, styles_shard) -> uint8 tensor [n,H,W,3]`` on this rank's
    block of frames and gather the full batch on every rank (frames in input order).

    contents: uint8 tensor [B,H,W,3]; styles: uint8 tensor [B or 1,Hs,Ws,3]."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = contents.shape[0]
    lo, hi = shard_range(B, world, rank)
    st = styles if styles.shape[0] == 1 else styles[lo:hi]
    local = process_fn(contents[lo:hi], st) if hi > lo else None
    if world == 1:
        return local
    sizes = [shard_range(B, world, r) for r in range(world)]
    maxn = max(h - l for l, h in sizes)
    ref = local if local is not None else None
    # every rank must know the output frame shape: take it from a rank that has frames
    shape = torch.tensor(list(ref.shape[1:]) if ref is not None else [0, 0, 0], dtype=torch.int64,
                         device=ref.device if ref is not None else contents.device)
    dist.all_reduce(shape, op=dist.ReduceOp.MAX, group=group)
    h, w, c = [int(v) for v in shape.tolist()]
    dev = ref.device if ref is not None else contents.device
    pad = torch.zeros((maxn, h, w, c), dtype=torch.uint8, device=dev)
    if ref is not None:
        pad[: hi - lo] = ref
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][: sizes[r][1] - sizes[r][0]] for r in range(world)], dim=0)
