"""GPU suite (-m gpu): the CUDA path, called through the C ABI (include/fps_b200.h), against
 (1) the committed golden vectors = outputs of the unmodified compiled reference,
 (2) the CPU oracle on the same seeded inputs (bit-exact: indices are integers),
 (3) at BASELINE.json's full sizes, the oracle where it finishes in seconds (its lazy kd-line form does)
     and otherwise the multi-threaded certifier (oracle_certify_fps) on a seeded subset of clouds plus
     size-independent properties on every cloud (range, first pick, distinctness, batch == single).
Nothing here reads /root/reference.
"""
import os
import sys
import threading

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from cases import CASES, input_sha, make_input  # noqa: E402

import fpsample_b200 as fps  # noqa: E402
from fpsample_b200 import capi, synth  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if capi.device_count() < 1:
        pytest.fail("no sm_100 device visible: the CUDA path cannot run and there is no CPU fallback")


def gpu_call(pc, call, p):
    if call == "vanilla":
        return capi.vanilla(pc, p["k"], p["start"])
    if call == "kdtree":
        return capi.kdtree(pc, p["k"], p["start"])
    if call == "npdu":
        return capi.npdu(pc, p["k"], p["w"], p["start"])
    if call == "npdukd":
        return capi.npdu_kdtree(pc, p["k"], p["w"], p["start"])
    return capi.kdline(pc, p["k"], p["h"], p["start"])


# ---- (1) golden vectors ------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_cuda_matches_reference_golden(case, golden):
    cid, spec, call, p = case
    pc = make_input(spec)
    assert input_sha(pc) == str(golden[cid + "__in"])
    before = capi.kernel_launches()
    got = gpu_call(pc, call, p)
    assert capi.kernel_launches() > before, "no CUDA kernel was launched"
    assert got.dtype == np.uint64 and got.shape == (p["k"],)
    np.testing.assert_array_equal(got, golden[cid].astype(np.uint64))


def test_python_api_is_a_drop_in(golden):
    """fpsample_b200.fps_sampling / bucket_fps_kdline_sampling: same call, same dtype, same indices."""
    np.random.seed(42)
    pc = np.random.rand(4096, 3)  # float64 in, like bench/test_bench.py:19-21
    out = fps.fps_sampling(pc, 1024, start_idx=0)
    assert out.dtype == np.uint64 and out.shape == (1024,)
    np.testing.assert_array_equal(out, golden["G0_vanilla"])
    np.testing.assert_array_equal(fps.bucket_fps_kdline_sampling(pc, 1024, 5, start_idx=0), golden["G0_kd_h5"])
    np.testing.assert_array_equal(fps.bucket_fps_kdline_sampling(pc, 1024, h=7, start_idx=0), golden["G0_kd_h7"])
    np.testing.assert_array_equal(fps.fps_sampling(np.asfortranarray(pc), 1024, 0), golden["G0_vanilla"])
    ms = fps.fps_sampling(synth.uniform(77, 4096, 3), 256, [1, 2, 50, 4000])
    np.testing.assert_array_equal(ms, golden["multi_start"])
    np.random.seed(7)                      # random start honours np.random.seed (README.md:89-92)
    a = fps.fps_sampling(pc, 16)
    np.random.seed(7)
    s = np.random.randint(0, 4096)
    assert a[0] == s


# ---- (2) oracle on seeded inputs: shapes that exercise every kernel plan -----------------------------------
VANILLA_SHAPES = [  # n, d, k, start
    (1, 3, 1, 0), (2, 3, 2, 1), (33, 3, 33, 5), (1000, 3, 100, 7), (4096, 6, 512, 5), (5000, 2, 300, 1),
    (3000, 8, 200, 0), (16384, 3, 1024, 3), (777, 1, 200, 4), (2048, 5, 100, 9), (1500, 7, 100, 9),
    (40000, 3, 300, 11), (3000, 12, 100, 2), (100000, 3, 500, 0), (100000, 6, 300, 0), (300000, 3, 200, 0),
    (250000, 8, 64, 1), (1 << 20, 3, 64, 12345),
]


@pytest.mark.parametrize("n,d,k,s", VANILLA_SHAPES)
def test_vanilla_vs_oracle(n, d, k, s, oracle):
    pc = synth.uniform(n + d, n, d)
    got = capi.vanilla(pc, k, s)
    np.testing.assert_array_equal(got, oracle.fps_vanilla(pc, k, s), err_msg=capi.last_plan())


KD_SHAPES = [  # n, d, k, h, start
    (2, 3, 2, 1, 0), (64, 3, 20, 2, 1), (64, 3, 64, 6, 63), (1000, 3, 100, 3, 7), (4096, 3, 1024, 7, 3),
    (4096, 6, 512, 5, 5), (5000, 2, 300, 4, 1), (3000, 8, 200, 6, 0), (777, 1, 200, 3, 4), (4096, 3, 200, 12, 0),
    (16384, 3, 700, 6, 11), (15500, 3, 300, 5, 2), (13000, 3, 400, 7, 1), (9000, 6, 300, 5, 4),
    (50000, 3, 4096, 7, 0), (100000, 3, 2000, 9, 0), (100000, 6, 1000, 9, 0), (300000, 4, 1000, 8, 9),
]


@pytest.mark.parametrize("n,d,k,h,s", KD_SHAPES)
def test_kdline_vs_oracle(n, d, k, h, s, oracle):
    pc = synth.uniform(n + d + h, n, d)
    got = capi.kdline(pc, k, h, s)
    np.testing.assert_array_equal(got, oracle.kdline(pc, k, h, s), err_msg=capi.last_plan())


@pytest.mark.parametrize("seed", range(4))
def test_tie_clouds(seed, oracle):
    """integer lattices: every tie rule and degenerate-split clamp (KDTreeBase.h:142-146)."""
    d = (1, 2, 3, 6)[seed]
    g = synth.grid_ties(seed, 3000, d, levels=4 + seed)
    np.testing.assert_array_equal(capi.vanilla(g, 700, [5, 1, 9]), oracle.fps_vanilla(g, 700, [5, 1, 9]))
    np.testing.assert_array_equal(capi.vanilla(g, 3000, 0), oracle.fps_vanilla(g, 3000, 0))
    for h in (1, 4, 8, 11):
        np.testing.assert_array_equal(capi.kdline(g, 700, h, seed), oracle.kdline(g, 700, h, seed), err_msg=f"h={h}")


@pytest.mark.parametrize("n,d,levels,k,h,s", [(60000, 3, 24, 3000, 6, 5), (40000, 2, 40, 2500, 8, 0), (30000, 6, 3, 1500, 7, 9),
                                              (70000, 1, 5000, 2000, 5, 1), (20000, 3, 2, 500, 10, 0)])
def test_async_cluster_path_on_tie_lattices(n, d, levels, k, h, s, oracle):
    """clouds too big for one SM go through the coordinator/worker cluster kernel (the planner now prefers the grouped
    grid sampler for them: switched off here): ties, duplicates and near-empty buckets must not disturb its
    upper-bound logic."""
    g = synth.grid_ties(n + h, n, d, levels=levels)
    with capi.tuning(group=0):
        got = capi.kdline(g, k, h, s)
    assert any(x in capi.last_plan() for x in ("kdline_async_kernel", "kdline_stream_kernel")), capi.last_plan()
    np.testing.assert_array_equal(got, oracle.kdline(g, k, h, s), err_msg=capi.last_plan())
    got = capi.kdline(g, k, h, s)   # and the default route
    np.testing.assert_array_equal(got, oracle.kdline(g, k, h, s), err_msg=capi.last_plan())


def test_async_batches_and_cluster_sizes(oracle):
    for B, n, k, h in [(3, 30000, 800, 7), (20, 20000, 400, 6), (80, 16384, 300, 7), (160, 14000, 200, 5)]:
        pcs = synth.uniform_batch(8000 + B, B, n, 3)
        st = (np.arange(B) * 7) % n
        with capi.tuning(group=0):
            got = capi.kdline_batch(pcs, k, h, st, devices=[0])
        assert any(x in capi.last_plan() for x in ("kdline_async_kernel", "kdline_warp", "kdline_stream")), capi.last_plan()
        want = np.stack([oracle.kdline(pcs[b], k, h, int(st[b])) for b in range(B)])
        np.testing.assert_array_equal(got, want, err_msg=capi.last_plan())


@pytest.mark.parametrize("knobs", [dict(warp_global_minb=1), dict(warp_global_minb=1, stream_warps=1), dict(warp_global_minb=1, stream_warps=2),
                                   dict(warp_lazy=0), dict(warp_tmem=0), dict(warp_hybrid=1)],
                         ids=["stream", "stream-1warp", "stream-2warps", "eager", "no-tmem", "hybrid"])
def test_alternative_samplers(knobs, oracle):
    """every kd-line sampler must give the reference's indices, not only the one the planner prefers: the streaming
    kernel over global memory with 1, 2 and 4 warps per cloud (big batches of big clouds), the eager variant of the on-chip
    warp kernel, its shared-memory-only placement and the shared-memory + TMEM hybrid."""
    first = next(iter(knobs))
    want_plan = {"warp_global_minb": "kdline_stream_kernel", "warp_lazy": "eager", "warp_tmem": "tmem 0", "warp_hybrid": "hybrid"}[first]
    big = first == "warp_global_minb"
    with capi.tuning(group=0, **knobs):   # the planner's default for medium clouds is the grouped grid sampler: off here
        if first == "warp_hybrid":   # only clouds between the shared-memory and the TMEM limit
            shapes = [(16384, 3, 700, 7, 11, "u"), (15500, 3, 300, 5, 2, "g"), (16000, 3, 400, 6, 0, "l")]
        elif big:
            shapes = [(30000, 3, 900, 7, 5, "u"), (20000, 6, 500, 6, 0, "u"), (40000, 2, 800, 5, 3, "g"), (50000, 1, 700, 5, 1, "g"),
                      (25000, 3, 600, 7, 2, "l"), (9000, 8, 300, 7, 4, "u"), (12345, 4, 1000, 7, 12344, "g"),
                      (20001, 3, 20001, 7, 17, "g"), (30011, 5, 500, 3, 0, "u"), (70001, 3, 400, 9, 70000, "u")]
        else:
            shapes = [(4096, 3, 1024, 5, 0, "u"), (3000, 6, 500, 5, 7, "u"), (5000, 2, 700, 7, 1, "g"), (4096, 3, 600, 6, 9, "l")]
        for n, d, k, h, s, gen in shapes:
            pc = {"u": lambda: synth.uniform(n + d, n, d), "g": lambda: synth.grid_ties(n, n, d, levels=37),
                  "l": lambda: synth.lidar(n, n)}[gen]()
            got = capi.kdline(pc, k, h, s)
            assert want_plan in capi.last_plan(), capi.last_plan()
            np.testing.assert_array_equal(got, oracle.kdline(pc, k, h, s), err_msg=capi.last_plan())
        if big:
            pcs = synth.uniform_batch(6100, 12, 20000, 3)
            got = capi.kdline_batch(pcs, 300, 7, np.arange(12) * 3, devices=[0])
            assert want_plan in capi.last_plan(), capi.last_plan()
            np.testing.assert_array_equal(got, np.stack([oracle.kdline(pcs[b], 300, 7, 3 * b) for b in range(12)]))


@pytest.mark.parametrize("n,d,k,h,s,gen", [(300000, 3, 5000, 9, 0, "u"), (270001, 3, 3000, 8, 77, "l"), (150000, 3, 150000, 7, 3, "g"),
                                           (65536, 2, 4000, 6, 1, "g"), (50000, 6, 2500, 7, 9, "u"), (9000, 1, 9000, 5, 0, "g"),
                                           (200000, 4, 2000, 10, 5, "u"), (40000, 8, 1200, 6, 2, "u")])
def test_grid_sampler(n, d, k, h, s, gen, oracle):
    """the whole-GPU sampler for one huge cloud (kdline_grid_kernel: points in shared memory, a batch of picks per
    grid-wide exchange) forced onto smaller clouds too: ties, duplicates (k = n drives every distance to 0), odd
    sizes, every padded dimension."""
    with capi.tuning(grid=1):
        pc = {"u": lambda: synth.uniform(n + d, n, d), "g": lambda: synth.grid_ties(n, n, d, levels=23),
              "l": lambda: synth.lidar(n, n)}[gen]()
        got = capi.kdline(pc, k, h, s)
        assert "kdline_grid_kernel" in capi.last_plan(), capi.last_plan()
        np.testing.assert_array_equal(got, oracle.kdline(pc, k, h, s), err_msg=capi.last_plan())
        if n <= 65536:   # a batch runs cloud after cloud in one launch
            pcs = np.stack([pc, pc[::-1].copy(), synth.uniform(n, n, d)])
            gb = capi.kdline_batch(pcs, min(k, 500), h, [s, 0, 1], devices=[0])
            assert "kdline_grid_kernel" in capi.last_plan(), capi.last_plan()
            for b in range(3):
                np.testing.assert_array_equal(gb[b], oracle.kdline(pcs[b], min(k, 500), h, [s, 0, 1][b]))


@pytest.mark.parametrize("n,d,k,s,gen", [(4096, 3, 1024, 0, "u"), (50000, 3, 3000, 17, "l"), (3000, 2, 3000, 5, "g"), (7777, 6, 900, 3, "u"),
                                         (20000, 1, 500, 0, "g"), (130000, 3, 1500, 9, "u"), (999, 8, 999, 998, "u")])
def test_kdtree_vs_oracle(n, d, k, s, gen, oracle):
    """bucket_fps_kdtree_sampling: the full kd permutation is built on the GPU (kdtree_build_kernel), the vanilla
    kernels sample the permuted rows, positions are mapped back to ids."""
    pc = {"u": lambda: synth.uniform(n + d, n, d), "g": lambda: synth.grid_ties(n, n, d, levels=11),
          "l": lambda: synth.lidar(n, n)}[gen]()
    got = capi.kdtree(pc, k, s)
    assert "kdtree_build_kernel" in capi.last_plan(), capi.last_plan()
    np.testing.assert_array_equal(got, oracle.kdtree(pc, k, s), err_msg=capi.last_plan())


def test_kdtree_batch_and_python_api(oracle):
    pcs = synth.uniform_batch(4400, 9, 3000, 3)
    st = (np.arange(9) * 131) % 3000
    got = fps.bucket_fps_kdtree_sampling_batch(pcs, 400, st, devices=[0])
    assert got.dtype == np.uint64 and got.shape == (9, 400)
    for b in range(9):
        np.testing.assert_array_equal(got[b], oracle.kdtree(pcs[b], 400, int(st[b])))
    one = fps.bucket_fps_kdtree_sampling(pcs[3].astype(np.float64), 400, start_idx=int(st[3]))
    np.testing.assert_array_equal(one, got[3])
    with pytest.raises(RuntimeError):   # the reference's dimension limit (src/wrapper.hpp:105-107)
        fps.bucket_fps_kdtree_sampling(synth.uniform(1, 100, 9), 10, start_idx=0)
    with pytest.raises(NotImplementedError):   # src/lib.cpp:482-485
        fps._bucket_fps_kdtree_sampling(pcs[0], 10, np.array([1, 2], dtype=np.uint64))


@pytest.mark.parametrize("n,d,k,h,s,gen,B", [(16384, 3, 4096, 7, 0, "u", 5), (30000, 3, 900, 7, 11, "u", 1), (20000, 6, 500, 6, 3, "u", 3),
                                             (40000, 2, 800, 5, 1, "g", 2), (60000, 3, 2000, 9, 0, "l", 1), (9000, 1, 9000, 5, 2, "g", 2),
                                             (12345, 4, 700, 8, 5, "u", 150), (25000, 8, 600, 7, 9, "u", 2), (100000, 3, 3000, 7, 4, "u", 2),
                                             (105000, 2, 1500, 9, 0, "g", 1)])
def test_group_sampler(n, d, k, h, s, gen, B, oracle):
    """batches of medium clouds on groups of CTAs (kdline_grid_kernel, flat mode: every warp publishes its own keys,
    slices = kd subtrees): more clouds than groups, ties, duplicates, every padded dimension."""
    with capi.tuning(group=1):
        mk = {"u": lambda i: synth.uniform(n + d + i, n, d), "g": lambda i: synth.grid_ties(n + i, n, d, levels=23),
              "l": lambda i: synth.lidar(n + i, n)}[gen]
        pcs = np.stack([mk(i) for i in range(B)])
        st = (np.arange(B) * 37 + s) % n
        got = capi.kdline_batch(pcs, k, h, st, devices=[0])
        assert "kdline_grid_kernel" in capi.last_plan() and "flat" in capi.last_plan(), capi.last_plan()
        for b in sorted({0, B // 2, B - 1}):
            np.testing.assert_array_equal(got[b], oracle.kdline(pcs[b], k, h, int(st[b])), err_msg=f"cloud {b}: {capi.last_plan()}")


@pytest.mark.parametrize("n,d,k,starts,gen", [(20000, 3, 700, [5, 1, 9], "g"), (16384, 2, 5000, [16383], "g"), (70000, 1, 2000, [0], "g"),
                                              (30000, 6, 1500, [7, 7, 29999], "u"), (40000, 8, 500, [3], "u"), (300000, 3, 3000, [11], "l"),
                                              (17000, 3, 17000, [4], "g")])
def test_vanilla_kd_route(n, d, k, starts, gen, oracle):
    """fps_sampling on big clouds runs the same exact recurrence pruned by a kd permutation (kdline_grid_kernel, IDS mode):
    ties still go to the highest ORIGINAL index, start lists are forced in order, k = n drives every distance to 0."""
    pc = {"u": lambda: synth.uniform(n + d, n, d), "g": lambda: synth.grid_ties(n, n, d, levels=17),
          "l": lambda: synth.lidar(n, n)}[gen]()
    got = capi.vanilla(pc, k, starts)
    assert "vanilla FPS via kd permutation" in capi.last_plan(), capi.last_plan()
    if n * k <= 3e8:
        np.testing.assert_array_equal(got, oracle.fps_vanilla(pc, k, starts), err_msg=capi.last_plan())
    else:
        ok, where = oracle.certify_vanilla(pc, got, n_forced=len(starts))
        assert ok, f"first bad round {where} ({capi.last_plan()})"
    with capi.tuning(vanilla_kd=0):   # and the brute-force kernels still agree
        np.testing.assert_array_equal(capi.vanilla(pc, min(k, 300), starts), got[:min(k, 300)])
        assert "kd permutation" not in capi.last_plan()


def test_vanilla_kd_batch_with_starts(oracle):
    pcs = synth.uniform_batch(9100, 7, 20000, 3)
    st = (np.arange(7) * 2999) % 20000
    got = capi.vanilla_batch(pcs, 600, st, devices=[0])
    assert "vanilla FPS via kd permutation" in capi.last_plan(), capi.last_plan()
    for b in range(7):
        np.testing.assert_array_equal(got[b], oracle.fps_vanilla(pcs[b], 600, int(st[b])))


def test_unaligned_and_strided_inputs(oracle):
    buf = synth.uniform(3, 4097 * 3 + 1, 1).ravel()
    pc = buf[1:1 + 4097 * 3].reshape(4097, 3)           # base address 4 bytes off any 16-byte boundary
    assert pc.ctypes.data % 16 != 0 or (pc.ctypes.data + 4) % 16 != 0
    np.testing.assert_array_equal(capi.vanilla(pc, 500, 0), oracle.fps_vanilla(pc, 500, 0))
    np.testing.assert_array_equal(capi.kdline(pc, 500, 5, 0), oracle.kdline(pc, 500, 5, 0))
    wide = synth.uniform(4, 3000, 8)
    view = wide[:, 2:5]                                  # non-contiguous view: the front-end copies (forcecast)
    np.testing.assert_array_equal(fps.fps_sampling(view, 300, 1), oracle.fps_vanilla(np.ascontiguousarray(view), 300, 1))


def test_kdline_build_matches_oracle(oracle):
    """perm / leaf ranges / tight boxes of the GPU build == the oracle's."""
    import torch
    # the small shapes go through kdsmall_kernel (the build behind the one-warp-per-cloud sampler): tie lattices in
    # D = 1, 2, 3 make every misplaced permutation slot visible, which a tie-free cloud's indices would hide
    for (n, d, h, gen) in [(4096, 3, 5, "u"), (20000, 3, 7, "l"), (3000, 6, 6, "g"), (100000, 3, 9, "u"), (64, 2, 6, "g"),
                           (3000, 1, 6, "g"), (3000, 2, 6, "g"), (5000, 2, 7, "g"), (3000, 3, 6, "g"), (4099, 3, 5, "u"),
                           (2000, 5, 4, "u"), (8000, 3, 7, "l"), (4096, 3, 8, "u"),
                           # one CTA of 1024 threads per cloud, index arrays in global memory (cfg 3's shape)
                           (16384, 3, 7, "u"), (16000, 2, 7, "g"), (12000, 4, 6, "u"), (17000, 3, 8, "g"),
                           # the edges of what one SM holds: 16-bit indices up to 65 535 points, 8 dimensions
                           (50000, 1, 8, "g"), (65535, 1, 5, "u"), (65536, 1, 5, "u"), (6000, 8, 6, "u"), (33, 3, 5, "u")]:
        pc = {"u": lambda: synth.uniform(n, n, d), "l": lambda: synth.lidar(n, n),
              "g": lambda: synth.grid_ties(n, n, d)}[gen]()
        S = 1 << h
        dp = torch.from_numpy(pc).cuda()
        perm = torch.empty(n, dtype=torch.int32, device="cuda")
        lo = torch.empty(S + 1, dtype=torch.int32, device="cuda")
        box = torch.empty(S * 2 * d, dtype=torch.float32, device="cuda")
        wsb = capi.workspace_bytes(capi.ALGO_KDLINE, 1, n, d, 1, h)
        ws = torch.empty(wsb + 256, dtype=torch.uint8, device="cuda")
        wp = (ws.data_ptr() + 255) & ~255
        capi.kdline_build_dev(dp.data_ptr(), 1, n, d, h, perm.data_ptr(), lo.data_ptr(), box.data_ptr(), wp, wsb,
                              torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        operm, obounds, obox = oracle.kdline_build(pc, h)
        np.testing.assert_array_equal(perm.cpu().numpy().astype(np.uint64), operm)
        glo = lo.cpu().numpy().astype(np.int64)
        gbox = box.cpu().numpy().reshape(S, 2, d)
        keep = np.flatnonzero(np.diff(glo) > 0)          # empty slots are allowed (early 'count==1' leaves)
        np.testing.assert_array_equal(glo[keep], obounds[:-1].astype(np.int64))
        assert glo[-1] == n
        np.testing.assert_array_equal(gbox[keep], obox)
    # a batch: several clouds per CTA through the dynamic scheduler, shared memory reused from cloud to cloud
    B, n, d, h = 500, 2500, 2, 6
    pcs = np.stack([synth.grid_ties(900 + b, n, d, levels=5 + b % 7) for b in range(B)])
    dp = torch.from_numpy(pcs).cuda()
    perm = torch.empty((B, n), dtype=torch.int32, device="cuda")
    wsb = capi.workspace_bytes(capi.ALGO_KDLINE, B, n, d, 1, h)
    ws = torch.empty(wsb + 256, dtype=torch.uint8, device="cuda")
    capi.kdline_build_dev(dp.data_ptr(), B, n, d, h, perm.data_ptr(), 0, 0, (ws.data_ptr() + 255) & ~255, wsb,
                          torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert "kdsmall_kernel" in capi.last_plan(), capi.last_plan()
    got = perm.cpu().numpy().astype(np.uint64)
    for b in range(0, B, 7):
        np.testing.assert_array_equal(got[b], oracle.kdline_build(pcs[b], h)[0], err_msg=f"cloud {b}")


def test_kdline_pick_counts_around_output_blocks(oracle):
    """the one-warp-per-cloud sampler writes ids in blocks of 32 picks, one pick late: every residue of k matters"""
    pc = synth.uniform(321, 4096, 3)
    pcs = synth.uniform_batch(322, 9, 3000, 3)
    for k in (1, 2, 31, 32, 33, 34, 63, 64, 65, 97, 993, 1025):
        np.testing.assert_array_equal(capi.kdline(pc, k, 5, 7), oracle.kdline(pc, k, 5, 7), err_msg=f"k={k}")
        got = capi.kdline_batch(pcs, k, 4, None, devices=[0])
        for b in range(9):
            np.testing.assert_array_equal(got[b], oracle.kdline(pcs[b], k, 4, 0), err_msg=f"k={k} cloud {b}")


def test_gpu_resident_arrays_through_the_python_api(oracle):
    """arrays that already live in HBM (anything with __cuda_array_interface__, here torch tensors)
    are sampled where they are -- same indices as the host-array call, for every entry point."""
    import torch
    pcs = synth.uniform_batch(4400, 12, 4096, 3)
    dp = torch.from_numpy(pcs).cuda()
    st = [int(x) for x in (np.arange(12) * 5) % 4096]
    before = capi.kernel_launches()
    np.testing.assert_array_equal(fps.bucket_fps_kdline_sampling_batch(dp, 300, 5, st), fps.bucket_fps_kdline_sampling_batch(pcs, 300, 5, st))
    np.testing.assert_array_equal(fps.fps_sampling_batch(dp, 200, 3), fps.fps_sampling_batch(pcs, 200, 3))
    np.testing.assert_array_equal(fps.bucket_fps_kdtree_sampling_batch(dp, 100), fps.bucket_fps_kdtree_sampling_batch(pcs, 100))
    assert capi.kernel_launches() > before
    one = dp[3]
    np.testing.assert_array_equal(fps.bucket_fps_kdline_sampling(one, 500, 5, start_idx=9), oracle.kdline(pcs[3], 500, 5, 9))
    np.testing.assert_array_equal(fps.fps_sampling(one, 500, start_idx=9), oracle.fps_vanilla(pcs[3], 500, 9))
    np.testing.assert_array_equal(fps.bucket_fps_kdtree_sampling(one, 64, start_idx=1), oracle.kdtree(pcs[3], 64, 1))
    # produced on the device right before the call, on torch's stream: the library waits for it
    fresh = dp * 2.0 + 1.0
    np.testing.assert_array_equal(fps.bucket_fps_kdline_sampling_batch(fresh, 64, 4), fps.bucket_fps_kdline_sampling_batch(fresh.cpu().numpy(), 64, 4))
    with pytest.raises(TypeError):
        fps.fps_sampling_batch(dp.double(), 10)                     # no implicit cast on the device
    with pytest.raises(TypeError):
        fps.fps_sampling_batch(dp.transpose(1, 2), 10)              # not C-contiguous
    big = torch.from_numpy(synth.uniform_batch(4500, 3, 20000, 3)).cuda()    # the grouped grid sampler reads it too
    np.testing.assert_array_equal(fps.bucket_fps_kdline_sampling_batch(big, 256, 7), fps.bucket_fps_kdline_sampling_batch(big.cpu().numpy(), 256, 7))


def test_sequential_sum_tiles_are_bit_exact():
    """the split value is a strictly sequential binary32 sum (KDTreeBase.h:151-158); the build kernels evaluate it tile by
    tile as an integer prefix scan (csrc/seqsum.cuh), and for long columns in two phases: the tiles are prepared in parallel
    under a guessed binade and one warp walks their records (tile = -512).  Bit-equal to numpy's sequential float32
    accumulate on columns that hit every branch: single-signed, zero-mean, exact ties, lattices, wild magnitudes, inf."""
    import torch
    g = np.random.default_rng(77)
    n = 300_000
    lid = synth.lidar(31, n)
    cols = {
        "uniform": synth.uniform(5, n, 3)[:, 0], "uniform-0.5": synth.uniform(6, n, 3)[:, 1] - np.float32(0.5),
        "lidar x": lid[:, 0], "lidar y": lid[:, 1], "lidar z": lid[:, 2], "sorted lidar x": np.sort(lid[:, 0]),
        "halves": (g.integers(-50, 50, n) * 0.5).astype(np.float32), "quarters": (g.integers(0, 64, n) * 0.25).astype(np.float32),
        "lattice": synth.grid_ties(3, n, 3)[:, 1], "gauss": (g.standard_normal(n) * 30).astype(np.float32),
        "mixed": np.where(np.arange(n) % 97 == 0, g.random(n) * 1e6, g.random(n) * 1e-3).astype(np.float32),
        "tiny": (g.random(n) * 1e-30).astype(np.float32), "huge": (g.random(n) * 1e30).astype(np.float32),
        "inf": np.where(np.arange(n) == 200_000, np.inf, g.random(n)).astype(np.float32),
        "short": synth.uniform(8, 700, 3)[:, 2], "one": np.array([3.25], np.float32),
    }
    out = torch.zeros(1, dtype=torch.float32, device="cuda")
    fast = torch.zeros(1, dtype=torch.int32, device="cuda")
    some_fast = 0
    for name, col in cols.items():
        col = np.ascontiguousarray(col, dtype=np.float32)
        want = np.add.accumulate(col, dtype=np.float32)[-1]           # strictly sequential
        d = torch.from_numpy(col).cuda()
        for tile in (256, 512, -512):
            capi.seqsum_dev(d.data_ptr(), col.size, out.data_ptr(), fast.data_ptr(), tile, torch.cuda.current_stream().cuda_stream)
            got = out.cpu().numpy()[0]
            assert got.tobytes() == want.tobytes(), f"{name} tile={tile}: {got!r} != {want!r}"
            some_fast += int(fast.cpu().numpy()[0])
            if tile == -512 and name in ("uniform", "quarters", "gauss"):
                assert int(fast.cpu().numpy()[0]) > 0.5 * (col.size // 512), f"{name}: the two-phase sum fell back to the chain on most tiles"
    assert some_fast > 1000   # the scan path really ran


def test_npdu_matches_the_oracle_and_the_front_end_is_a_drop_in(oracle, golden):
    """fps_npdu_sampling: the index-window heuristic, bit-identical to src/lib.cpp:272-340"""
    for n, d, k, w, s, gen in [(4096, 3, 1024, 64, 0, "u"), (3000, 2, 500, 10, 7, "g"), (20000, 3, 2048, 156, 11, "l"),
                               (100000, 3, 4096, 390, 5, "u"), (777, 1, 300, 33, 776, "g"), (2000, 6, 300, 1, 5, "u"),
                               (5000, 3, 5000, 4999, 3, "u"), (300, 3, 1, 8, 299, "u"), (257, 3, 257, 300, 0, "g")]:
        pc = {"u": lambda: synth.uniform(n + d, n, d), "g": lambda: synth.grid_ties(n, n, d), "l": lambda: synth.lidar(n, n)}[gen]()
        np.testing.assert_array_equal(capi.npdu(pc, k, w, s), oracle.fps_npdu(pc, k, w, s), err_msg=str((n, d, k, w, s, gen)))
    pcs = synth.uniform_batch(4600, 300, 4096, 3)
    st = (np.arange(300) * 13) % 4096
    got = capi.npdu_batch(pcs, 512, 128, st, devices=[0])
    for b in range(0, 300, 17):
        np.testing.assert_array_equal(got[b], oracle.fps_npdu(pcs[b], 512, 128, int(st[b])))
    np.random.seed(42)
    pc = np.random.rand(4096, 3)
    out = fps.fps_npdu_sampling(pc, 1024, start_idx=0)                 # default window: n / n_samples * 16 = 64
    assert out.dtype == np.uint64 and out.shape == (1024,)
    np.testing.assert_array_equal(out, golden["G0_npdu"])
    with pytest.warns(UserWarning):
        np.testing.assert_array_equal(fps.fps_npdu_sampling(pc, 100, w=10**6, start_idx=3), oracle.fps_npdu(pc, 100, 4095, 3))


def test_npdu_kdtree_matches_the_oracle_and_the_front_end_is_a_drop_in(oracle, golden):
    """fps_npdu_kdtree_sampling (second half): the k-nearest-neighbour heuristic of src/lib.cpp:369-465.
    The GPU finds the k nearest by brute force + radix select; the oracle by brute force + sort; both are pinned to the
    compiled reference (nanoflann) by the golden cases.  Clouds in general position only (see include/fps_b200.h on ties)."""
    for n, d, k, w, s, gen in [(4096, 3, 1024, 64, 0, "u"), (3000, 6, 500, 33, 7, "u"), (20000, 3, 700, 300, 5, "l"), (1000, 1, 1000, 17, 999, "u"),
                               (50000, 3, 300, 2000, 1, "u"), (2500, 12, 400, 2499, 3, "u"), (600, 2, 100, 600, 0, "u"), (70000, 3, 200, 70, 9, "u")]:
        pc = {"u": lambda: synth.uniform(n + d, n, d), "l": lambda: synth.lidar(n, n)}[gen]()
        got = capi.npdu_kdtree(pc, k, w, s)
        assert "npdu_knn_kernel" in capi.last_plan(), capi.last_plan()
        np.testing.assert_array_equal(got, oracle.fps_npdu_kdtree(pc, k, w, s), err_msg=str((n, d, k, w, s, gen)))
    pcs = synth.uniform_batch(8800, 40, 5000, 3)
    st = (np.arange(40) * 97) % 5000
    got = capi.npdu_kdtree_batch(pcs, 256, 80, st, devices=[0])
    for b in range(0, 40, 3):
        np.testing.assert_array_equal(got[b], oracle.fps_npdu_kdtree(pcs[b], 256, 80, int(st[b])))
    np.random.seed(42)
    pc = np.random.rand(4096, 3)
    out = fps.fps_npdu_kdtree_sampling(pc, 1024, start_idx=0)          # default: n / n_samples * 16 = 64 neighbours
    assert out.dtype == np.uint64 and out.shape == (1024,)
    np.testing.assert_array_equal(out, golden["G0_npdukd"])
    with pytest.warns(UserWarning, match="k is too large"):            # src/fpsample/__init__.py:136-138
        np.testing.assert_array_equal(fps.fps_npdu_kdtree_sampling(pc, 100, w=10**6, start_idx=3), oracle.fps_npdu_kdtree(pc, 100, 4096, 3))
    with pytest.raises(NotImplementedError):                            # src/lib.cpp:385-390
        fps._fps_npdu_kdtree_sampling(pc.astype(np.float32), 10, 5, np.array([1, 2], dtype=np.uint64))


def test_device_pointer_entries(oracle):
    import torch
    B, n, d, k, h = 6, 5000, 3, 400, 5
    pcs = synth.uniform_batch(70, B, n, d)
    dp = torch.from_numpy(pcs).cuda()
    st = torch.arange(B, dtype=torch.int64, device="cuda")
    out = torch.empty((B, k), dtype=torch.int64, device="cuda")
    stream = torch.cuda.Stream()
    for algo in (capi.ALGO_VANILLA, capi.ALGO_KDLINE):
        wsb = capi.workspace_bytes(algo, B, n, d, k, h)
        ws = torch.empty(wsb + 256, dtype=torch.uint8, device="cuda")
        wp = (ws.data_ptr() + 255) & ~255
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            if algo == capi.ALGO_VANILLA:
                capi.vanilla_batch_dev(dp.data_ptr(), B, n, d, k, st.data_ptr(), out.data_ptr(), wp, wsb, stream.cuda_stream)
            else:
                capi.kdline_batch_dev(dp.data_ptr(), B, n, d, k, st.data_ptr(), h, out.data_ptr(), wp, wsb, stream.cuda_stream)
        stream.synchronize()
        got = out.cpu().numpy().astype(np.uint64)
        fn = (lambda b: oracle.fps_vanilla(pcs[b], k, b)) if algo == capi.ALGO_VANILLA else (lambda b: oracle.kdline(pcs[b], k, h, b))
        np.testing.assert_array_equal(got, np.stack([fn(b) for b in range(B)]))
    with pytest.raises(capi.FpsError) as e:               # too-small workspace is refused, not overrun
        capi.kdline_batch_dev(dp.data_ptr(), B, n, d, k, 0, h, out.data_ptr(), wp, 16, 0)
    assert e.value.rc == 5


def test_concurrent_callers(oracle):
    """the module advertises free-threading (src/lib.cpp:581): concurrent calls must not interfere."""
    pcs = [synth.uniform(900 + i, 3000 + 17 * i, 3) for i in range(8)]
    res = [None] * 8

    def work(i):
        res[i] = (fps.fps_sampling(pcs[i], 300, i), fps.bucket_fps_kdline_sampling(pcs[i], 300, 4, i))

    th = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    [t.start() for t in th]
    [t.join() for t in th]
    for i in range(8):
        np.testing.assert_array_equal(res[i][0], oracle.fps_vanilla(pcs[i], 300, i))
        np.testing.assert_array_equal(res[i][1], oracle.kdline(pcs[i], 300, 4, i))


# ---- (3) BASELINE.json configs at full size ---------------------------------------------------------------
def test_cfg1_vanilla_4096(oracle):
    pc = synth.uniform(1, 4096, 3)
    np.testing.assert_array_equal(capi.vanilla(pc, 1024, 0), oracle.fps_vanilla(pc, 1024, 0))


def test_cfg2_batch_1024x4096_h5(oracle):
    pcs = synth.uniform_batch(1000, 1024, 4096, 3)
    got = capi.kdline_batch(pcs, 1024, 5)
    assert got.shape == (1024, 1024) and got.dtype == np.uint64
    want = np.stack([oracle.kdline(pcs[b], 1024, 5, 0) for b in range(1024)])
    np.testing.assert_array_equal(got, want)
    # "indices checked bit-exact vs vanilla" (cfg 2) can only mean: exact FPS over the permuted array from
    # position 0.  Without float ties that IS vanilla started at the same point:
    van = capi.vanilla_batch(pcs, 1024, got[:, 0].copy())
    same = (van == got).all(axis=1)
    assert same.mean() > 0.95                             # the ~1.5 % that differ are exact float ties (F5)
    for b in np.flatnonzero(~same)[:8]:
        assert oracle.certify_vanilla(pcs[b], van[b])[0]  # vanilla's own tie rule holds on those
    sub = np.arange(0, 1024, 64)
    np.testing.assert_array_equal(van[sub], np.stack([oracle.fps_vanilla(pcs[b], 1024, int(got[b, 0])) for b in sub]))


def test_cfg3_batch_64x16384_h7(oracle):
    pcs = synth.uniform_batch(2000, 64, 16384, 3)
    got = capi.kdline_batch(pcs, 4096, 7)
    np.testing.assert_array_equal(got, np.stack([oracle.kdline(pcs[b], 4096, 7, 0) for b in range(64)]))
    van = capi.vanilla_batch(pcs, 4096)
    for b in range(0, 64, 8):
        assert oracle.certify_vanilla(pcs[b], van[b])[0], b
    assert (van[:, 0] == 0).all()


@pytest.mark.parametrize("gen", ["uniform", "lidar"])
def test_cfg4_single_1m_to_64k_h9(gen, oracle, golden):
    pc = synth.uniform(5, 2**20, 3) if gen == "uniform" else synth.lidar(6, 2**20)
    got = capi.kdline(pc, 65536, 9, 0)
    assert "kdline_grid_kernel" in capi.last_plan(), capi.last_plan()
    np.testing.assert_array_equal(got, golden["G5_kd_h9" if gen == "uniform" else "G6_kd_h9"])
    assert len(np.unique(got)) == 65536


def test_cfg4_vanilla_1m_certified(oracle):
    """vanilla at 2^20 points: the sequential oracle needs minutes, the threaded certifier seconds."""
    pc = synth.uniform(5, 2**20, 3)
    k = 8192
    got = capi.vanilla(pc, k, 0)
    ok, where = oracle.certify_vanilla(pc, got)
    assert ok, f"first bad round {where} ({capi.last_plan()})"


@pytest.mark.parametrize("d", [3, 6])
def test_cfg5_batch_100k_to_8192(d, oracle, golden):
    """full single-cloud size, a 64-cloud slice of the 4096-cloud batch (the full batch is bench.py's job);
    kd-line h=7 against the oracle for every cloud, vanilla certified on a subset."""
    B = 64
    pcs = synth.uniform_batch(3000, B, 100000, d)
    got = capi.kdline_batch(pcs, 8192, 7)
    np.testing.assert_array_equal(got[0], golden["cfg5_b0_d3_kd" if d == 3 else "cfg5_b0_d6_kd"])
    for b in range(B):
        assert len(np.unique(got[b])) == 8192 and got[b].max() < 100000
    for b in range(0, B, 4 if d == 3 else 16):
        np.testing.assert_array_equal(got[b], oracle.kdline(pcs[b], 8192, 7, 0), err_msg=f"cloud {b}")
    van = capi.vanilla_batch(pcs[:16], 8192)
    if d == 3:
        np.testing.assert_array_equal(van[1], golden["cfg5_b1_d3_vanilla"])
    for b in (0, 7, 15):
        assert oracle.certify_vanilla(pcs[b], van[b])[0], b


@pytest.mark.parametrize("d,B,ncheck", [(3, 1200, 12), (6, 320, 4)])
def test_cfg5_production_plan(d, B, ncheck, oracle):
    """BASELINE.json configs[4] on the plan the full 4096-cloud batch (and every shard of it down to one eighth) takes:
    the grid-wide build + the streaming sampler (teams of 1 / 2 / 4 warps per cloud by batch size, 16 warps per SM).  A
    seeded subset of the clouds is checked against the oracle, every cloud for the properties an exact FPS result has."""
    n, k, h = 100000, 8192, 7
    pcs = synth.uniform_batch(3000, B, n, d)
    got = capi.kdline_batch(pcs, k, h, devices=[0])
    plan = capi.last_plan()
    # 1200 clouds of 3-D points = one full wave of two-warp teams (8 per SM) + a tail on four-warp teams; 6-D records: four warps
    assert "kdline_stream_kernel" in plan and "gb_* grid-wide build" in plan and "WPC=4" in plan, plan
    assert ("1184 clouds x WPC=2" in plan and "16 clouds x WPC=4" in plan) if d == 3 else "WPC=2" not in plan, plan
    for b in range(B):
        assert got[b].max() < n
    for b in range(0, B, 7):
        assert len(np.unique(got[b])) == k, b
    rng = np.random.default_rng(d)
    for b in sorted({0, B - 1} | set(int(x) for x in rng.integers(0, B, ncheck - 2))):
        np.testing.assert_array_equal(got[b], oracle.kdline(pcs[b], k, h, 0), err_msg=f"cloud {b}: {plan}")
    # the shard one of eight GPUs gets takes the same kernel with 4 warps per cloud; the full batch's one-warp teams are
    # forced onto a slice of it too (a knob: they buy nothing over two-warp teams)
    if d == 3:
        shard = capi.kdline_batch(pcs[:512], k, h, devices=[0])
        assert "kdline_stream_kernel" in capi.last_plan() and "WPC=4" in capi.last_plan(), capi.last_plan()
        np.testing.assert_array_equal(shard, got[:512])
        with capi.tuning(stream_warps=1):
            one = capi.kdline_batch(pcs[:400], k, h, devices=[0])
            assert "WPC=1" in capi.last_plan(), capi.last_plan()
        np.testing.assert_array_equal(one, got[:400])
        with capi.tuning(stream_split=0):   # one team size for the whole batch (the planner before the wave split)
            two = capi.kdline_batch(pcs[600:], k, h, devices=[0])
            assert "600 clouds x WPC=4" in capi.last_plan(), capi.last_plan()
        np.testing.assert_array_equal(two, got[600:])


def test_executed_work_counters_of_the_streaming_sampler(oracle):
    """COUNT=1: the kernel counts what it executes (bench.py's W_exec).  The reference's lazy scheme is reproduced pass for
    pass as long as no pending list fills up, so points scanned / point-updates equal the oracle's counters."""
    n, d, k, h, B = 30000, 3, 900, 7, 6
    pcs = synth.uniform_batch(7700, B, n, d)
    with capi.tuning(group=0, warp_global_minb=1, count=1):
        got = capi.kdline_batch(pcs, k, h, devices=[0])
        assert "kdline_stream_kernel" in capi.last_plan(), capi.last_plan()
        cnt = capi.debug_counters(capi.DBG_STREAM)
    pts, pu, passes, early, tests, picks, clouds, stored = [int(x) for x in cnt[:8]]
    assert 0 < stored <= pts
    assert clouds == B and picks == B * (k - 1) and tests == picks * 2**h
    want_pu = 0
    for b in range(B):
        np.testing.assert_array_equal(got[b], oracle.kdline(pcs[b], k, h, 0))
        # the reference also applies the LAST pick (KDLineTree.h:77-85), whose update nobody observes: the kernel applies
        # samples 0 .. k-2, which is the oracle's work for k - 1 picks
        want_pu += oracle.kdline(pcs[b], k - 1, h, 0, return_stats=True)[1]["point_updates"]
    assert early == 0, "R=16 pending samples per bucket should never fill up on uniform 3-D clouds"
    assert pu == want_pu, (pu, want_pu)
    assert pts <= pu


def test_batch_equals_single_and_per_cloud_starts(oracle):
    pcs = synth.uniform_batch(4000, 37, 4096, 3)
    st = np.arange(37) * 11
    got = capi.vanilla_batch(pcs, 256, st)
    np.testing.assert_array_equal(got, np.stack([capi.vanilla(pcs[b], 256, int(st[b])) for b in range(37)]))
    np.testing.assert_array_equal(got, np.stack([oracle.fps_vanilla(pcs[b], 256, int(st[b])) for b in range(37)]))
    gk = fps.bucket_fps_kdline_sampling_batch(pcs, 256, 5, start_idx=list(st))
    np.testing.assert_array_equal(gk, np.stack([oracle.kdline(pcs[b], 256, 5, int(st[b])) for b in range(37)]))


def test_multi_device_sharding_in_process(oracle):
    nd = capi.device_count()
    if nd < 2:
        pytest.skip("one GPU visible")
    pcs = synth.uniform_batch(5000, 4 * nd + 1, 4096, 3)
    a = capi.kdline_batch(pcs, 512, 5, devices=list(range(nd)))
    b = capi.kdline_batch(pcs, 512, 5, devices=[0])
    np.testing.assert_array_equal(a, b)
