"""One-process-per-GPU plumbing for the batched entries.

Clouds are independent, so a batch is cut into contiguous shards (remainder to the low ranks -- the same rule the
in-process multi-device path of the C ABI uses, csrc/capi.cu run_batch) and every rank samples its shard on its own GPU
with NO data-path collective.  The only exchange is the final gather of the index arrays to rank 0, and it lives in the C
layer (csrc/comm.cu): uint32 on the wire, one group of ncclSend / ncclRecv over NVLink, widened on rank 0's GPU.

This module only carries the 128-byte NCCL id from rank 0 to the other ranks -- through the launcher's rendezvous
(MASTER_ADDR / MASTER_PORT, a plain TCP exchange; no torch) -- and offers the same gather over an initialised
torch.distributed group ("torch" transport) for hosts without GPUs: that is what the world_size-2 gloo tests on CPU use.
"""
from __future__ import annotations

import os
import socket
import struct
import time
from typing import Callable, Optional, Tuple

import numpy as np

_ID_BYTES = 128
_PORT_OFFSET = 1   # the id exchange listens next to the launcher's own store
_comm_rank: Optional[int] = None   # this process's rank in the C layer's communicator (init_comm)


def shard_range(n_clouds: int, world: int, rank: int) -> Tuple[int, int]:
    """-> (first cloud, number of clouds) of `rank`'s contiguous shard."""
    assert world >= 1 and 0 <= rank < world
    base, rem = divmod(n_clouds, world)
    nb = base + (1 if rank < rem else 0)
    b0 = rank * base + min(rank, rem)
    return b0, nb


def env_rank_world() -> Tuple[int, int, int]:
    """(rank, world, local rank) as torchrun exports them; (0, 1, 0) outside a launcher."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def exchange_bytes(payload: Optional[bytes], rank: int, world: int, addr: Optional[str] = None, port: Optional[int] = None,
                   timeout: float = 120.0) -> bytes:
    """Rank 0's `payload` to every rank over TCP (rank 0 listens on MASTER_ADDR : MASTER_PORT + 1).  Pure stdlib."""
    if world == 1:
        assert payload is not None
        return payload
    addr = addr or os.environ.get("MASTER_ADDR", "127.0.0.1")
    port = port if port is not None else int(os.environ.get("MASTER_PORT", "29500")) + _PORT_OFFSET
    if rank == 0:
        assert payload is not None
        srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        srv.bind((addr, port))
        srv.listen(world)
        srv.settimeout(timeout)
        try:
            for _ in range(world - 1):
                c, _peer = srv.accept()
                with c:
                    c.sendall(struct.pack("<I", len(payload)) + payload)
        finally:
            srv.close()
        return payload
    deadline = time.time() + timeout
    while True:
        try:
            with socket.create_connection((addr, port), timeout=5.0) as c:
                c.settimeout(timeout)
                hdr = _recv_exact(c, 4)
                return _recv_exact(c, struct.unpack("<I", hdr)[0])
        except (ConnectionRefusedError, socket.timeout, OSError):
            if time.time() > deadline:
                raise TimeoutError(f"rank {rank}: no id from rank 0 at {addr}:{port}")
            time.sleep(0.05)


def _recv_exact(c: socket.socket, n: int) -> bytes:
    buf = b""
    while len(buf) < n:
        part = c.recv(n - len(buf))
        if not part:
            raise ConnectionError("peer closed the id exchange")
        buf += part
    return buf


def init_comm(rank: Optional[int] = None, world: Optional[int] = None, addr: Optional[str] = None, port: Optional[int] = None) -> None:
    """Create the C layer's NCCL communicator on the CURRENT CUDA device of this process (one process per GPU).
    Collective over all ranks.  The id comes from fps_b200_comm_unique_id on rank 0."""
    from . import capi
    r, w, _ = env_rank_world()
    rank = r if rank is None else rank
    world = w if world is None else world
    uid = capi.comm_unique_id() if rank == 0 else None
    uid = exchange_bytes(uid, rank, world, addr, port)
    assert len(uid) == _ID_BYTES
    capi.comm_init(uid, world, rank)
    global _comm_rank
    _comm_rank = rank


def gather_indices(local: np.ndarray, n_clouds: int, group=None, device=None, transport: Optional[str] = None) -> Optional[np.ndarray]:
    """Gather every rank's [nb, k] uint64 index block to rank 0 -> [n_clouds, k] uint64 (None elsewhere).

    transport "nccl" (default when init_comm ran): the C layer's gather (csrc/comm.cu).  "torch": the same exchange over an
    initialised torch.distributed group (gloo on hosts without GPUs)."""
    from . import capi
    if transport is None:
        transport = "nccl" if capi.comm_ranks() > 0 else "torch"
    if transport == "nccl":
        return capi.gather_indices(local, n_clouds, is_root=(_comm_rank == 0))
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
                   group=None, device=None, transport: Optional[str] = None) -> Optional[np.ndarray]:
    """Run `sample_fn` (e.g. a partial of fps_sampling_batch bound to this rank's device) on this rank's
    shard and gather the indices to rank 0."""
    return gather_indices(sample_fn(pcs_local), n_clouds, group=group, device=device, transport=transport)


def bucket_fps_kdline_sampling_sharded(pcs_local, n_clouds: int, n_samples: int, h: int, start_idx=None) -> Optional[np.ndarray]:
    """Fused form (needs init_comm): this rank's shard is sampled on its GPU, the indices never leave the device until rank 0
    copies all of them to the host once.  -> [n_clouds, n_samples] uint64 on rank 0, None elsewhere."""
    from . import capi
    return capi.kdline_batch_sharded(pcs_local, n_clouds, n_samples, h, start_idx, is_root=(_comm_rank == 0))


def fps_sampling_sharded(pcs_local, n_clouds: int, n_samples: int, start_idx=None) -> Optional[np.ndarray]:
    from . import capi
    return capi.vanilla_batch_sharded(pcs_local, n_clouds, n_samples, start_idx, is_root=(_comm_rank == 0))
