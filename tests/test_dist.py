"""The N>1 plumbing.  CPU suite: contiguous sharding, the TCP exchange that carries the NCCL id, and the index gather to rank 0
on world_size-2 gloo (the compute stand-in is the oracle: tests may use it as a checker / stand-in, the product never does).
GPU suite (needs >= 2 devices, skipped otherwise): the C layer's NCCL gather, one process per GPU and in one process."""
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
        out = D.sample_sharded(fn, pcs, B, transport="torch")
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


def _id_worker(rank, world, port, q):
    try:
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        payload = bytes(range(128)) if rank == 0 else None
        q.put((rank, D.exchange_bytes(payload, rank, world, "127.0.0.1", port, timeout=60)))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))


def test_id_exchange_over_tcp():
    """the 128-byte NCCL id reaches every rank through a plain TCP exchange on the launcher's address (no torch)"""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_id_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs[1:] + procs[:1]:   # the clients may come up before the server
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=30)
    assert all(res[r] == bytes(range(128)) for r in range(3)), res


def test_comm_entry_points_fail_loudly_without_a_communicator():
    from fpsample_b200 import capi
    assert capi.comm_ranks() == 0
    with pytest.raises(capi.FpsError) as e:
        capi.gather_indices(np.zeros((2, 4), dtype=np.uint64), 2)
    assert e.value.rc == 7   # FPS_ERR_NCCL


def test_binding_nccl_first_does_not_break_a_later_torch_import():
    """The library binds NCCL at run time.  A process must never hold two different libnccl.so.2 (the second user is handed
    the first one's symbols: `undefined symbol: ncclDevCommCreate` when the system's older NCCL was bound before PyTorch was
    imported), so the python package names the pip-installed wheel PyTorch itself would load (fps_b200_nccl_library)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("from fpsample_b200 import capi\n"
            "v = capi.nccl_version()\n"          # binds NCCL before torch is in the process
            "import torch\n"
            "t = torch.cuda.nccl.version() if hasattr(torch.cuda, 'nccl') else None\n"
            "print(v, t, capi._bundled_nccl())\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    v, rest = r.stdout.split(None, 1)
    if "None" not in rest.split()[-1]:   # there is a bundled wheel: the versions agree
        assert int(v) > 0, r.stdout


def _nccl_worker(rank, world, port, B, n, k, h, q):
    try:
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import ctypes
        from fpsample_b200 import capi
        import torch
        torch.cuda.set_device(rank)
        D.init_comm(rank, world, "127.0.0.1", port)
        b0, nb = D.shard_range(B, world, rank)
        pcs = np.stack([synth.uniform(700 + b, n, 3) for b in range(b0, b0 + nb)])
        fused = D.bucket_fps_kdline_sampling_sharded(pcs, B, k, h)                       # indices stay on the device until rank 0 copies them
        sep = D.gather_indices(capi.kdline_batch(pcs, k, h, devices=[rank]), B)          # sampled separately, gathered afterwards
        van = D.fps_sampling_sharded(pcs, B, k)
        capi.comm_destroy()
        q.put((rank, None if fused is None else (np.array(fused), np.array(sep), np.array(van))))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))


@pytest.mark.gpu
@pytest.mark.parametrize("B", [5, 16])
def test_nccl_gather_one_process_per_gpu(B, oracle):
    from fpsample_b200 import capi
    if capi.device_count() < 2:
        pytest.skip("one GPU visible")
    import multiprocessing as mp
    world = min(capi.device_count(), 4)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n, k, h = 20000, 300, 6
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, port, B, n, k, h, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert not isinstance(res[0], str), res[0]
    assert all(res[r] is None for r in range(1, world)), res
    fused, sep, van = res[0]
    pcs = [synth.uniform(700 + b, n, 3) for b in range(B)]
    np.testing.assert_array_equal(fused, np.stack([oracle.kdline(pc, k, h, 0) for pc in pcs]))
    np.testing.assert_array_equal(sep, fused)
    np.testing.assert_array_equal(van, np.stack([oracle.fps_vanilla(pc, k, 0) for pc in pcs]))


@pytest.mark.gpu
def test_nccl_gather_in_one_process(oracle):
    """SURVEY.md 8(e), single process: ncclCommInitAll, one communicator + stream per device, the whole batch in host memory"""
    from fpsample_b200 import capi
    nd = capi.device_count()
    if nd < 2:
        pytest.skip("one GPU visible")
    B, n, k, h = 4 * nd + 1, 4096, 512, 5
    pcs = synth.uniform_batch(5000, B, n, 3)
    capi.comm_init_local(list(range(nd)))
    try:
        assert capi.comm_ranks() == nd
        got = capi.kdline_batch_sharded(pcs, B, k, h)
        np.testing.assert_array_equal(got, capi.kdline_batch(pcs, k, h, devices=[0]))
        np.testing.assert_array_equal(got[B - 1], oracle.kdline(pcs[B - 1], k, h, 0))
    finally:
        capi.comm_destroy()
