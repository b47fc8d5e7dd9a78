"""Multi-GPU plumbing: photons (sample indices) shard across ranks, films add up.

SURVEY.md §8(e): rank ``r`` of ``R`` renders samples ``[r*spp/R, (r+1)*spp/R)`` of EVERY pixel.
Philox counters are the global path id ``(i*W+j)*spp + s`` so the ranks' streams are disjoint by
construction and the union over ranks is bit-identical to what one rank would have drawn.  The
only exchange of a pass is ONE sum-reduction of the ``H*W*3`` float64 film (``torch.distributed``:
NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests).  The reference has no counterpart — it
is single-process (``numba.prange`` threads only, R/src/path_tracing_old.py:142-149).
"""
from __future__ import annotations


def sample_range(rank: int, world_size: int, spp: int) -> tuple[int, int]:
    """Contiguous, gap-free, balanced split of ``range(spp)``."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return rank * spp // world_size, (rank + 1) * spp // world_size


def reduce_film(film, group=None, root: int | None = None):
    """Sum ``film`` (a torch tensor, CUDA for NCCL / CPU for gloo) over the ranks, in place.
    ``root=None`` → all-reduce (every rank gets the image); else reduce to ``root``."""
    import torch.distributed as dist
    if root is None:
        dist.all_reduce(film, op=dist.ReduceOp.SUM, group=group)
    else:
        dist.reduce(film, dst=root, op=dist.ReduceOp.SUM, group=group)
    return film
