"""Ray sharding across one-process-per-GPU ranks.

The scene (triangles, BVH, materials) is replicated on every GPU; rank g traces global ray ids
[g*N/G, (g+1)*N/G) of EVERY context (the Philox stream is keyed by (seed, context, ray id), so the
union over ranks is the same set of paths whatever G is).  Each rank owns a full-size partial
histogram.  The only exchange is at the end: ONE sum-reduce of the histograms to rank 0 plus a min /
max reduce of the two track-range words, then rank 0 finalises (x 1/N, direct sound, x gain^2,
src/Scene.cpp:286-316).  Works with any torch.distributed backend (NCCL over NVLink on the GPU box,
gloo in the CPU tests)."""
from __future__ import annotations

from typing import Tuple


def shard_bounds(n_rays: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, exhaustive, non-overlapping ray-id ranges; sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return n_rays * rank // world, n_rays * (rank + 1) // world


def reduce_partials(hist, first_sample, real_length, dst: int = 0) -> None:
    """In-place reduce to `dst`: hist by SUM, first_sample by MIN, real_length by MAX.
    `hist` [tracks, bins] float32; `first_sample` / `real_length` [tracks] int32 (separate, contiguous)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    dist.reduce(hist, dst=dst, op=dist.ReduceOp.SUM)
    dist.reduce(first_sample, dst=dst, op=dist.ReduceOp.MIN)
    dist.reduce(real_length, dst=dst, op=dist.ReduceOp.MAX)
