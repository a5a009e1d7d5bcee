"""Multi-GPU sharding of a batch of independent QPs.

The path has no data-path collective: every state's cascade is independent (the reference solves exactly one
per tick, ref:src/ForceAcc.cpp:189, ref:src/QPPVMPlugin.cpp:246).  Rank r of G owns the contiguous block
[r*B/G, (r+1)*B/G).  torch.distributed (NCCL over NVLink on GPUs, gloo on CPU for the tests) is used only to
scatter records from a root rank and to gather outputs back; the kernel writes its outputs directly into the
tensor that is the gather's send buffer (no staging copy).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist


def partition(batch: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block split; the first (batch % world) ranks take one extra problem.  Returns (start, count)."""
    base, extra = divmod(batch, world)
    count = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, count


def scatter_records(records_root: Optional[torch.Tensor], batch: int, rec_doubles: int, device, root: int = 0) -> torch.Tensor:
    """Root holds (batch, rec_doubles) float64 on `device`; every rank returns its own (count, rec_doubles) block."""
    world, rank = dist.get_world_size(), dist.get_rank()
    start, count = partition(batch, world, rank)
    mine = torch.empty((count, rec_doubles), dtype=torch.float64, device=device)
    ops = []
    if rank == root:
        for r in range(world):
            s, c = partition(batch, world, r)
            if r == root:
                mine.copy_(records_root[s:s + c])
            elif c:
                ops.append(dist.P2POp(dist.isend, records_root[s:s + c].contiguous(), r))
    elif count:
        ops.append(dist.P2POp(dist.irecv, mine, root))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return mine


def gather_outputs(out_local: torch.Tensor, batch: int, root: int = 0) -> Optional[torch.Tensor]:
    """Inverse of scatter_records for the (count, out_doubles) outputs; returns the full tensor on root, else None."""
    world, rank = dist.get_world_size(), dist.get_rank()
    ops, full = [], None
    if rank == root:
        full = torch.empty((batch, out_local.shape[1]), dtype=out_local.dtype, device=out_local.device)
        for r in range(world):
            s, c = partition(batch, world, r)
            if r == root:
                full[s:s + c].copy_(out_local)
            elif c:
                ops.append(dist.P2POp(dist.irecv, full[s:s + c], r))
    elif out_local.shape[0]:
        ops.append(dist.P2POp(dist.isend, out_local.contiguous(), root))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return full


def solve_sharded(solve: Callable[[torch.Tensor], torch.Tensor], records_root: Optional[torch.Tensor], batch: int,
                  rec_doubles: int, device, root: int = 0) -> Optional[torch.Tensor]:
    """scatter -> per-rank solve (``solve(records) -> outputs``, e.g. ``lambda r: solver.solve_batch(r)[0]``) -> gather."""
    mine = scatter_records(records_root, batch, rec_doubles, device, root)
    out = solve(mine)
    return gather_outputs(out, batch, root)
