"""Batch sharding for multi-GPU inference: every video is independent end to end (no cross-sample
op anywhere in reference src/modelling/models.py:16-195), so rank r of W simply takes a contiguous
slice of the batch; weights are replicated and the data path needs NO collective. Logits are
gathered only where a caller wants them on one rank (evaluation), via torch.distributed.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

BATCHED_KEYS = ("categories", "boxes", "scores", "frame_types", "lengths", "labels",
                "src_key_padding_mask_boxes", "src_key_padding_mask_frames")


def shard_bounds(batch_size: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of rank ``rank``; the first ``batch_size % world_size`` ranks get one more."""
    if not 0 <= rank < world_size:
        raise ValueError("rank outside [0, world_size)")
    base, rem = divmod(batch_size, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch: Dict, world_size: int, rank: int) -> Dict:
    B = batch["categories"].shape[0]
    lo, hi = shard_bounds(B, world_size, rank)
    out = {}
    for k, v in batch.items():
        if isinstance(v, torch.Tensor) and k in BATCHED_KEYS:
            out[k] = v[lo:hi]
        elif isinstance(v, (list, tuple)) and len(v) == B:
            out[k] = v[lo:hi]
        else:
            out[k] = v
    return out


def gather_logits(local_logits: torch.Tensor, batch_size: int) -> torch.Tensor:
    """All-gathers per-rank logits back into batch order (evaluation helper; not on the hot path)."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return local_logits
    world = dist.get_world_size()
    sizes = [shard_bounds(batch_size, world, r) for r in range(world)]
    max_rows = max(hi - lo for lo, hi in sizes)
    padded = local_logits.new_zeros((max_rows, local_logits.shape[1]))
    padded[: local_logits.shape[0]] = local_logits
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)
