"""Multi-GPU partitioning of the hot path: independent frame pairs / camera streams are sharded
across ranks with NO data-path collective; one process per GPU, one FlowEngine per
process.  torch.distributed is used only to gather the small per-unit results (the velocity scalars
the nodes publish) and for barriers/timing in bench.py.

The reference itself is single-GPU (``.devcontainer/docker-compose.yml:22-28`` reserves one GPU), so
this is new design rather than a mirror of reference code.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np


def shard_indices(n_units: int, rank: int, world: int) -> List[int]:
    """Round-robin ownership: unit i (frame pair or camera stream) lives on rank i mod world, so a
    stream's temporal state (previous frame, LK points) always stays on the same GPU."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world %d/%d" % (rank, world))
    return list(range(rank, n_units, world))


def owner_of(unit: int, world: int) -> int:
    return unit % world


def gather_unit_values(local: Dict[int, float], n_units: int, group=None) -> np.ndarray:
    """All ranks contribute {unit index: scalar}; every rank receives the dense float64 [n_units]
    array (NaN where no rank reported).  Uses one all_reduce(SUM) over a masked buffer — works on
    gloo (CPU tests) and NCCL (GPU) alike."""
    import torch
    import torch.distributed as dist

    vals = torch.zeros(n_units, dtype=torch.float64)
    have = torch.zeros(n_units, dtype=torch.float64)
    for k, v in local.items():
        vals[k] = float(v)
        have[k] = 1.0
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        backend = dist.get_backend(group)
        buf = torch.stack([vals, have])
        if backend == "nccl":
            buf = buf.cuda()
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        buf = buf.cpu()
        vals, have = buf[0], buf[1]
    out = vals.numpy().copy()
    out[have.numpy() == 0] = np.nan
    if (have.numpy() > 1).any():
        raise RuntimeError("a unit was reported by more than one rank")
    return out


def split_batches(units: Sequence[int], batch: int) -> List[List[int]]:
    """Chunks of at most `batch` units (one batched engine call each)."""
    return [list(units[i:i + batch]) for i in range(0, len(units), batch)]
