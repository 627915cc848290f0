"""Multi-GPU use of the eigensolve path: independent replicas, one process per GPU.

A single factorisation is one dependency tree and does not split into independent units
(SURVEY.md section 8e); what does shard without any data-path collective are the reference's
sweeps: the Reynolds / shift sweep of `.examples/eigenvalues.py:61-108` (one (A, sigma) per
iteration, same sparsity pattern) and the direct + adjoint pair of
`Sensitivity/__init__.py:158-311`.  Tasks are dealt round-robin to ranks; the only communication
is the final gather of the (tiny) results over `torch.distributed` (NCCL on GPUs, gloo on CPU).
"""

from __future__ import annotations

from typing import Any, Callable, Sequence


def shard_tasks(n_tasks: int, rank: int, world_size: int) -> list[int]:
    """Indices of the tasks owned by `rank` (round-robin, so similar-cost neighbours spread out)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    return list(range(rank, n_tasks, world_size))


def run_sharded(tasks: Sequence[Any], fn: Callable[[Any], Any], *, rank: int | None = None,
                world_size: int | None = None, gather: bool = True) -> list[Any]:
    """Run `fn(task)` for this rank's share; with `gather`, every rank returns all results in task
    order (all_gather_object), otherwise only its own (None elsewhere)."""
    import torch.distributed as dist

    if rank is None or world_size is None:
        if dist.is_available() and dist.is_initialized():
            rank, world_size = dist.get_rank(), dist.get_world_size()
        else:
            rank, world_size = 0, 1
    mine = shard_tasks(len(tasks), rank, world_size)
    local = {i: fn(tasks[i]) for i in mine}
    out: list[Any] = [None] * len(tasks)
    if gather and world_size > 1:
        parts: list[Any] = [None] * world_size
        dist.all_gather_object(parts, local)
        for part in parts:
            for i, v in part.items():
                out[i] = v
    else:
        for i, v in local.items():
            out[i] = v
    return out
