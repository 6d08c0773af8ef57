"""Host-side plumbing of the multi-GPU path: one process per GPU (torchrun), torch.distributed only carries
the NCCL unique id and small host objects; the data path (panel broadcasts) is NCCL inside libgpp_b200.so.

Block rows of NB rows are dealt cyclically: block b lives on rank b % world at local block b // world."""
from __future__ import annotations

import numpy as np


def block_owner(b, world):
    return b % world


def local_row_map(M, NB, rank, world):
    """Global row indices of the rows rank owns, in local storage order."""
    rows = []
    b = rank
    while b * NB < M:
        rows.append(np.arange(b * NB, min(M, (b + 1) * NB)))
        b += world
    return np.concatenate(rows) if rows else np.zeros(0, dtype=np.int64)


def assemble_from_locals(pieces, M, NB):
    """Rebuild the dense M x M matrix from per-rank row pieces (list indexed by rank)."""
    world = len(pieces)
    out = np.zeros((M, M))
    for r, piece in enumerate(pieces):
        idx = local_row_map(M, NB, r, world)
        assert piece.shape == (idx.size, M), (piece.shape, idx.size, M)
        out[idx] = piece
    return out


def init_engine_distributed(engine, dist):
    """Create the NCCL communicator of `engine` from an initialised torch.distributed group."""
    from . import _lib
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [_lib.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    engine.dist_init(rank, world, box[0])
    return rank, world
