"""Multi-GPU film sharding: one process per GPU, the scene replicated,
the film sharded by interleaved row strips, ONE sum-reduce to assemble it.

Every pixel is owned by exactly one rank, which renders all of its samples in ascending order, so
each film value is computed by a single rank exactly as a 1-GPU render computes it; the other
ranks contribute exact zeros, and x + 0 is exact -- the assembled film is bit-identical to the
single-GPU film (up to the sign of zero).  There is no exchange inside the path loop, so the
only collective is the final reduce (NCCL over NVLink on the B200 box, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def owned_rows(height: int, strip_rows: int, world: int, rank: int) -> list[int]:
    """Film rows (row 0 = image top) owned by `rank`: strips k with k % world == rank."""
    if world <= 1 or strip_rows <= 0:
        return list(range(height))
    return [r for r in range(height) if (r // strip_rows) % world == rank]


def render_sharded(render_fn, height: int, width: int, strip_rows: int = 8, device: str | torch.device = "cpu",
                   dst: int = 0) -> torch.Tensor:
    """Renders this rank's strips with `render_fn(film, strip_rows, world, rank)` -- which must fill
    the owned rows of the zero-initialised (3, H, W, 3) tensor `film` (colour, normal, albedo) --
    and sum-reduces the films onto rank `dst`.  Returns the film (complete on `dst`)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    film = torch.zeros((3, height, width, 3), dtype=torch.float32, device=device)
    render_fn(film, strip_rows, world, rank)
    if world > 1:
        dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM)
    return film
