"""One factorisation / eigensolve split over the GPUs of a node (stage 1).

The reference reaches several ranks through PETSc / SLEPc / MUMPS on `PETSc.COMM_WORLD`
(`Solver/utils.py:196-203`; "fully MPI-parallelized", `README.md:43`).  Here: one process per GPU (torchrun),
`torch.distributed` for the rendezvous only; the data path is NCCL called from the CUDA library:

* the assembly tree is cut by proportional mapping (`csrc/partition.cpp`): whole sub-trees go to single GPUs, the
  fronts above the cut are replicated (every GPU factors and sweeps them redundantly on identical data);
* factorisation: the contribution blocks of the sub-tree roots are broadcast from their owners (one NCCL group);
* every operator application: ONE all-reduce over the replicated rows carries the SpMV partial sums of those rows
  and the sub-tree roots' contribution vectors; the Krylov basis is row-sharded (sub-tree rows per GPU, replicated
  rows kept in step), each Gram-Schmidt pass costs one small all-reduce (coefficients and |w|^2 in one payload).

Usage (every rank, same calls in the same order):

    es = EigenSolver(A, M, cfg)                       # A, M: the full matrices on every rank
    es.solver.set_backend_options(partition="auto")   # rank / world from torch.distributed
    pairs = es.solve()                                # identical, complete eigenpairs on every rank
"""

from __future__ import annotations

from . import _lib

__all__ = ["world_info", "attach_comm", "make_handle"]


def world_info(group=None) -> tuple[int, int]:
    """(rank, world) of the initialised torch.distributed group, (0, 1) without one."""
    try:
        import torch.distributed as dist
    except Exception:
        return 0, 1
    if not (dist.is_available() and dist.is_initialized()):
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def attach_comm(handle: "_lib.Handle", group=None) -> None:
    """Create the handle's NCCL communicator: rank 0 draws the unique id, torch.distributed hands it round
    (collective: every rank of the group calls this at the same point)."""
    import torch  # noqa: F401  (loads libnccl.so.2 into the process; the CUDA library dlopens it by name)
    import torch.distributed as dist

    rank, world = world_info(group)
    if world != handle.world or rank != handle.rank:
        raise ValueError(f"handle was created for rank {handle.rank} / {handle.world}, the process group says {rank} / {world}")
    ids = [_lib.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    handle.set_comm(ids[0])


def make_handle(n: int, device: int, group=None) -> "_lib.Handle":
    """A handle for this rank's part of a partitioned solve (plain handle when there is one rank)."""
    rank, world = world_info(group)
    return _lib.Handle(n, device, rank, world)
