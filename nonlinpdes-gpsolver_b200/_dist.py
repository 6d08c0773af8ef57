"""Host-side plumbing of the multi-GPU path: one process per GPU (torchrun), torch.distributed only carries the NCCL
unique id and small host objects; the data path (panel gathers, diagonal-block broadcasts) is NCCL inside libgpp_b200.so.

Ownership: NB x NB block (bi, bc) of a replicated matrix belongs to rank (bi % P) * Q + (bc % Q) of a P x Q process grid
(2-D block-cyclic; the default Q = 1 is block-row cyclic)."""
from __future__ import annotations

import numpy as np


def block_owner(bi, bc, P, Q=1):
    return (bi % P) * Q + (bc % Q)


def owned_blocks(nblk, rank, P, Q=1, lower=True):
    """(bi, bc) pairs of an nblk x nblk block matrix that `rank` owns (lower triangle incl. diagonal by default)."""
    p, q = divmod(rank, Q)
    return [(bi, bc) for bi in range(p, nblk, P) for bc in range(q, (bi + 1) if lower else nblk, Q)]


def held_rows(M, NB, rank, P, Q=1):
    """Global row indices of the block rows a rank assembles (bi % P == p), ascending."""
    p = rank // Q
    rows = [np.arange(b * NB, min(M, (b + 1) * NB)) for b in range(p, (M + NB - 1) // NB, P)]
    return np.concatenate(rows) if rows else np.zeros(0, dtype=np.int64)


def grid_shape(world, Q=None):
    """Default process grid: P x 1 (multi-GPU); Q may be forced (GPP_DIST_Q)."""
    import os
    if Q is None:
        Q = int(os.environ.get("GPP_DIST_Q", "1"))
    if Q < 1 or world % Q:
        raise ValueError(f"Q = {Q} does not divide the number of ranks {world}")
    return world // Q, Q


def init_engine_distributed(engine, dist, Q=None):
    """Create the NCCL communicator of `engine` from an initialised torch.distributed group."""
    from . import _lib
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [_lib.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    engine.dist_init(rank, world, box[0])
    P, Qv = grid_shape(world, Q)
    engine.dist_set_grid(P, Qv)
    return rank, world
