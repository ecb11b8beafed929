"""One-process-per-GPU plumbing for the batched entries.

Clouds are independent, so a batch is cut into contiguous shards (remainder to the low ranks -- the same
rule the in-process multi-device path of the C ABI uses, csrc/capi.cu run_batch) and every rank samples its
shard on its own GPU with NO data-path collective.  The only exchange is the final gather of the index
arrays to rank 0: uint32 on the wire (indices < 2^32), NCCL over NVLink when the process group is NCCL,
gloo in the CPU tests.  torch / torch.distributed are plumbing here, never compute.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def shard_range(n_clouds: int, world: int, rank: int) -> Tuple[int, int]:
    """-> (first cloud, number of clouds) of `rank`'s contiguous shard."""
    assert world >= 1 and 0 <= rank < world
    base, rem = divmod(n_clouds, world)
    nb = base + (1 if rank < rem else 0)
    b0 = rank * base + min(rank, rem)
    return b0, nb


def gather_indices(local: np.ndarray, n_clouds: int, group=None, device=None) -> Optional[np.ndarray]:
    """Gather every rank's [nb, k] uint64 index block to rank 0 -> [n_clouds, k] uint64 (None elsewhere)."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    k = local.shape[1]
    b0, nb = shard_range(n_clouds, world, rank)
    assert local.shape[0] == nb, f"rank {rank} holds {local.shape[0]} clouds, its shard is {nb}"
    nb_max = shard_range(n_clouds, world, 0)[1]
    wire = np.zeros((nb_max, k), dtype=np.uint32)
    wire[:nb] = local  # indices < 2^32 (checked by the C ABI: n < 0xfffffff0)
    t = torch.from_numpy(wire.view(np.int32))
    if device is not None:
        t = t.to(device, non_blocking=True)
    bufs = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, bufs, dst=0, group=group)
    if rank != 0:
        return None
    out = np.empty((n_clouds, k), dtype=np.uint64)
    for r in range(world):
        r0, rn = shard_range(n_clouds, world, r)
        out[r0:r0 + rn] = bufs[r][:rn].cpu().numpy().view(np.uint32)
    return out


def sample_sharded(sample_fn: Callable[[np.ndarray], np.ndarray], pcs_local: np.ndarray, n_clouds: int,
                   group=None, device=None) -> Optional[np.ndarray]:
    """Run `sample_fn` (e.g. a partial of fps_sampling_batch bound to this rank's device) on this rank's
    shard and gather the indices to rank 0."""
    return gather_indices(sample_fn(pcs_local), n_clouds, group=group, device=device)
