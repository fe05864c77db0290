"""Instance sharding over the GPUs of one box: contiguous blocks, no collective on the data path.

One process per GPU (torchrun): rank r of G solves instances [lo, hi) = shard_range(total, r, G) on its own
device with ``LCQProblemBatch.setInstanceOffset(lo)`` (the perturbStep generator is keyed by the GLOBAL instance
index, so sharding does not change any instance's arithmetic).  The only communication is after the solve:
a gather of the results to rank 0 and max/sum reductions of scalars -- through ``torch.distributed`` (NCCL on
the GPU box, gloo in the CPU test-suite).  Inside ONE process the C++ class LCQPow::LCQProblemBatch does the same
over a list of devices with a host-side gather.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of rank ``rank``: [total*rank//world, total*(rank+1)//world)."""
    if world <= 0 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard arguments")
    return total * rank // world, total * (rank + 1) // world


def shard_counts(total: int, world: int) -> list:
    return [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]


def _world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def gather_rows(local: np.ndarray, total: int, device: Optional[torch.device] = None) -> Optional[np.ndarray]:
    """Gather the row blocks of all ranks in rank order.  Every rank calls it; rank 0 receives the (total, ...)
    array, the others None.  ``device``: where the exchange buffers live (cuda for NCCL, cpu for gloo)."""
    world = _world()
    local = np.ascontiguousarray(local)
    if world == 1:
        return local
    rank = dist.get_rank()
    counts = shard_counts(total, world)
    assert local.shape[0] == counts[rank], (local.shape, counts, rank)
    width = int(np.prod(local.shape[1:])) if local.ndim > 1 else 1
    raw = np.zeros((max(counts), width * local.dtype.itemsize), dtype=np.uint8)
    raw[: counts[rank]] = local.reshape(counts[rank], -1).view(np.uint8)
    dev = device if device is not None else torch.device("cpu")
    mine = torch.from_numpy(raw).to(dev)
    bufs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(bufs, mine)   # blocks are small next to the solve; all_gather works on both backends
    if rank != 0:
        return None
    out = np.empty((total,) + local.shape[1:], dtype=local.dtype)
    lo = 0
    for r in range(world):
        blk = bufs[r].cpu().numpy()[: counts[r]].view(local.dtype).reshape((counts[r],) + local.shape[1:])
        out[lo: lo + counts[r]] = blk
        lo += counts[r]
    return out


def reduce_scalars(values: Sequence[float], op: str, device: Optional[torch.device] = None) -> list:
    """max / sum of a few scalars over the ranks (timings are reported as the max over ranks)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device if device is not None else torch.device("cpu"))
    if _world() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return [float(v) for v in t.tolist()]
