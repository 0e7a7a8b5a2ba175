"""Sample-split data parallelism: one process per GPU, one reduce per frame.

The reference is single-GPU (no NCCL/MPI call anywhere).  The path shards
naturally (reference kernel.cu:131-153: a thread touches only its own pixel),
and with the counter-based stream keyed on the GLOBAL sample index the union of
the ranks' sample ranges is exactly the 1-GPU sample set.  So: replicate the
scene (each rank uploads the same host description), rank k of G renders all
pixels for samples [k*spp/G, (k+1)*spp/G), and the fp32 accumulators are summed
with a single reduce to rank 0 (NCCL over NVLink on GPUs; gloo in CPU tests).
"""
from __future__ import annotations

from typing import Tuple


def sample_range(rank: int, world: int, spp: int) -> Tuple[int, int]:
    """[begin, end) of global sample indices for `rank`; ranges tile [0, spp) exactly."""
    if not (0 <= rank < world) or spp < 0:
        raise ValueError("bad rank/world/spp")
    return (rank * spp) // world, ((rank + 1) * spp) // world


def reduce_accumulators(accum, dst: int = 0):
    """Sum the per-rank accumulation buffers onto `dst` (torch tensor, in place)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum
