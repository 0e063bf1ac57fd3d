"""Stream sharding across the GPUs of one node: independent audio streams are split
contiguously over ranks, weights are replicated, and there is NO collective on the data path.
`torch.distributed` is used only by callers that want the shards gathered on one rank (or for the
timing barrier in bench.py)."""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_streams: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous split: rank r gets streams [start, end); the first n % world ranks get one extra."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(n_streams, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_sizes(n_streams: int, world_size: int) -> List[int]:
    return [shard_range(n_streams, r, world_size)[1] - shard_range(n_streams, r, world_size)[0] for r in range(world_size)]


def gather_rank_rows(vals, device=None) -> List[List[float]]:
    """Every rank contributes the same number of floats and receives all of them as [world][len(vals)]: one all_gather of
    a float64 tensor.  bench.py uses it for the per-GPU report ("per-GPU roofline report + aggregate");
    it is not on the data path.  Without an initialised process group the result is the caller's own row."""
    t = torch.tensor([float(v) for v in vals], dtype=torch.float64, device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [[float(x) for x in t]]
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, t)
    return [[float(x) for x in p.cpu()] for p in parts]


def enhance_sharded(enhance_fn: Callable[[torch.Tensor], torch.Tensor], audio: torch.Tensor,
                    rank: Optional[int] = None, world_size: Optional[int] = None, gather_to: Optional[int] = None):
    """Runs `enhance_fn` on this rank's shard of `audio` [B, T] (every rank passes the same global batch, or any
    tensor whose rows [start, end) are valid).  Returns the local result, or -- when `gather_to` is given -- the
    full [B, T'] result on that rank (None elsewhere).  The gather is the only communication and is optional."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    b = audio.shape[0]
    s, e = shard_range(b, rank, world_size)
    local = enhance_fn(audio[s:e]) if e > s else audio.new_zeros((0, audio.shape[1]))
    if gather_to is None or world_size == 1:
        return local
    sizes = shard_sizes(b, world_size)
    t_out = torch.tensor([local.shape[1] if local.numel() else 0], dtype=torch.int64)
    dist.all_reduce(t_out, op=dist.ReduceOp.MAX)
    width = int(t_out.item())
    if local.shape[0] == 0:
        local = audio.new_zeros((0, width))
    local = local.contiguous().cpu()
    if rank == gather_to:
        parts = [torch.empty((n, width), dtype=local.dtype) for n in sizes]
        # gloo / nccl gather needs equal sizes: pad every shard to the largest one
        m = max(sizes)
        bufs = [torch.empty((m, width), dtype=local.dtype) for _ in sizes]
        pad = torch.zeros((m, width), dtype=local.dtype)
        pad[: local.shape[0]] = local
        dist.gather(pad, bufs, dst=gather_to)
        for i, n in enumerate(sizes):
            parts[i] = bufs[i][:n]
        return torch.cat(parts, dim=0)
    m = max(sizes)
    pad = torch.zeros((m, width), dtype=local.dtype)
    pad[: local.shape[0]] = local
    dist.gather(pad, None, dst=gather_to)
    return None
