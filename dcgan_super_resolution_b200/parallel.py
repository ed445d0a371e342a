"""Data-parallel plumbing: one process per GPU, the minibatch sharded in contiguous
row blocks, parameters / Adam state replicated.  torch.distributed carries the rendezvous (the NCCL
unique id of the library's own communicator); the gradient / BN / loss all-reduces run inside
libdcgansr.so on its own stream.  The reference has no counterpart (single GPU, train.lua:169)."""
from __future__ import annotations

import os

import numpy as np


def env_rank():
    """(rank, local_rank, world_size) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)))


def shard_bounds(global_batch: int, world_size: int, rank: int):
    """Contiguous row block [lo, hi) of this rank; equal shards are required (the criteria divide by
    the global element count, train.lua:220 sizeAverage)."""
    if global_batch % world_size:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world_size}")
    per = global_batch // world_size
    return rank * per, (rank + 1) * per


def shard_batch(real: np.ndarray, world_size: int, rank: int) -> np.ndarray:
    lo, hi = shard_bounds(real.shape[0], world_size, rank)
    return real[lo:hi]


def exchange_unique_id(ctx, dist, device=None) -> bytes:
    """Rank 0 creates the NCCL unique id of the library communicator; everybody receives it through
    torch.distributed (any backend) and joins."""
    import torch
    rank = dist.get_rank()
    uid = ctx.comm_unique_id() if rank == 0 else bytes(128)
    t = torch.tensor(list(uid), dtype=torch.uint8, device=device if device is not None else "cpu")
    dist.broadcast(t, src=0)
    uid = bytes(t.cpu().tolist())
    ctx.comm_init(uid)
    return uid
