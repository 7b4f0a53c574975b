"""Multi-GPU plumbing: one process per GPU, independent shards, one tiny exchange at the end
(SURVEY.md 8e).  Works with torch.distributed's nccl (GPU) and gloo (CPU tests) backends."""
from __future__ import annotations

import math

import numpy as np


def shard_range(n: int, rank: int, world: int):
    """Contiguous block partition of range(n): ranks [0, n % world) get one extra item."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _isless(a, b):
    """Julia isless on Float64: NaN maximal, -0.0 < +0.0."""
    if math.isnan(a):
        return False
    if math.isnan(b):
        return True
    if a == 0.0 and b == 0.0:
        return math.copysign(1.0, a) < math.copysign(1.0, b)
    return a < b


def reduce_pairs(pairs):
    """Deterministic arg-max over (value, global index) pairs: Julia argmax semantics, lowest index wins ties;
    pairs with index < 0 (empty shards) are ignored."""
    best = None
    for v, i in pairs:
        if i < 0:
            continue
        if best is None or _isless(best[0], v) or (not _isless(v, best[0]) and i < best[1]):
            best = (v, i)
    return best if best is not None else (float("-inf"), -1)


def allgather_argmax(value: float, global_index: int, device=None):
    """NCCL has no arg-max op: all-gather the 16-byte (value, index) pairs and reduce locally on every rank."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value, global_index
    mine = torch.tensor([value, float(global_index)], dtype=torch.float64, device=device)
    out = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(out, mine)
    pairs = [(float(t[0]), int(t[1])) for t in torch.stack(out).cpu()]
    return reduce_pairs(pairs)


def sharded_argmax(score_shard, M: int, device=None):
    """score_shard(lo, hi) -> (local_best_idx in [0, hi-lo), value) on this rank's candidate block."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_range(M, rank, world)
    if hi > lo:
        idx, val = score_shard(lo, hi)
        gi = lo + idx
    else:
        val, gi = float("-inf"), -1
    return allgather_argmax(val, gi, device)


def sharded_loglik(loglik_shard, S: int, device=None):
    """loglik_shard(lo, hi) -> np.ndarray of hi-lo log-likelihoods; returns all S on every rank (all-gather)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_range(S, rank, world)
    mine = np.asarray(loglik_shard(lo, hi), dtype=np.float64) if hi > lo else np.empty(0)
    if world == 1:
        return mine
    width = -(-S // world)
    buf = torch.full((width,), float("nan"), dtype=torch.float64, device=device)
    buf[: hi - lo] = torch.as_tensor(mine, dtype=torch.float64, device=device)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    res = np.empty(S)
    for r in range(world):
        a, b = shard_range(S, r, world)
        res[a:b] = out[r][: b - a].cpu().numpy()
    return res
