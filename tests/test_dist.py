"""CPU suite: the N>1 plumbing (contiguous sharding + index gather to rank 0) on world_size-2 gloo.
The compute stand-in is the oracle (tests may use it as a checker / stand-in; the product never does)."""
import os
import socket
import sys

import numpy as np
import pytest

from fpsample_b200 import dist as D
from fpsample_b200 import synth


def test_shard_range_partitions_like_the_c_abi():
    for B in (1, 2, 7, 64, 1024, 4097):
        for W in (1, 2, 3, 4, 8):
            spans = [D.shard_range(B, W, r) for r in range(W)]
            assert spans[0][0] == 0 and sum(nb for _, nb in spans) == B
            for (a0, an), (b0, _) in zip(spans, spans[1:]):
                assert a0 + an == b0
            sizes = [nb for _, nb in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, n, k, h, q):
    try:
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import torch.distributed as dist

        from oracle import oracle as O
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        b0, nb = D.shard_range(B, world, rank)
        pcs = np.stack([synth.uniform(500 + b, n, 3) for b in range(b0, b0 + nb)])
        fn = lambda x: np.stack([O.kdline(c, k, h, 0) for c in x])
        out = D.sample_sharded(fn, pcs, B)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, None if out is None else out))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))


@pytest.mark.parametrize("B", [5, 8])
def test_two_rank_gloo_gather(B, oracle):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n, k, h = 512, 64, 3
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, n, k, h, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res[1] is None
    assert not isinstance(res[0], str), res[0]
    want = np.stack([oracle.kdline(synth.uniform(500 + b, n, 3), k, h, 0) for b in range(B)])
    assert res[0].dtype == np.uint64
    np.testing.assert_array_equal(res[0], want)
