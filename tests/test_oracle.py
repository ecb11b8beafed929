"""CPU suite: pins the oracle (oracle/fps_oracle.c) to the reference.

 1. against the committed golden vectors (outputs of the unmodified compiled reference,
    tests/golden/make_golden.py) -- runs everywhere, including the GPU box where /root/reference is absent;
 2. against the compiled reference itself (oracle/_ref, built by oracle/build_ref.sh) on fresh seeded inputs
    -- runs wherever the prebuilt oracle/_ref travelled to;
 3. internal consistency: lazy bucket form == eager form, certifier accepts / rejects.
"""
import hashlib
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from cases import CASES, input_sha, make_input  # noqa: E402

from fpsample_b200 import synth  # noqa: E402


def run_oracle(O, pc, call, p):
    if call == "vanilla":
        return O.fps_vanilla(pc, p["k"], p["start"])
    if call == "kdtree":
        return O.kdtree(pc, p["k"], p["start"])
    if call == "npdu":
        return O.fps_npdu(pc, p["k"], p["w"], p["start"])
    if call == "npdukd":
        return O.fps_npdu_kdtree(pc, p["k"], p["w"], p["start"])
    return O.kdline(pc, p["k"], p["h"], p["start"])


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_golden(case, oracle, golden):
    cid, spec, call, p = case
    pc = make_input(spec)
    assert input_sha(pc) == str(golden[cid + "__in"]), "seeded input generator drifted from the golden run"
    got = run_oracle(oracle, pc, call, p)
    assert got.dtype == np.uint64 and got.shape == (p["k"],)
    np.testing.assert_array_equal(got, golden[cid].astype(np.uint64))


def test_survey_known_answer_hashes(golden):
    """sha256 of the reference output (detects a mis-built reference / stale fixtures)."""
    H = lambda a: hashlib.sha256(np.asarray(a).astype(np.uint64).tobytes()).hexdigest()[:16]
    want = {"G0_vanilla": "e1d132502f57a5b3", "G0_kd_h3": "c63be3c65e118849", "G0_kd_h5": "355746f6503ce6c4",
            "G0_kd_h7": "0ef3532ac4c01cf7", "G1_vanilla": "02aef47c729706fe", "G1_kd_h5": "ba107f50b86e54d0",
            "G2_vanilla": "3e9952878357b1db", "G2_kd_h7": "c4b71366f5e34558", "G3_vanilla": "0258f2e7d4ae830a",
            "G3_kd_h7": "2e23baa12c1ca398", "G4_vanilla": "2389ed33b38dd428", "G4_kd_h7": "bc7d544ed8dcf0f8",
            "G5_kd_h9": "65867fc4b5baccc0"}
    for cid, h in want.items():
        assert H(golden[cid]) == h, cid


@pytest.fixture(scope="module")
def ref(oracle):
    r = oracle.load_reference()
    if r is None:
        pytest.skip("compiled reference (oracle/_ref) not present")
    return r


@pytest.mark.parametrize("seed", range(6))
def test_oracle_matches_compiled_reference_random(seed, oracle, ref):
    g = np.random.default_rng(100 + seed)
    n = int(g.integers(64, 6000))
    d = int(g.integers(1, 9))
    k = int(g.integers(1, n // 2))
    h = int(g.integers(1, 8))
    while 2**h > n:
        h -= 1
    s = int(g.integers(0, n))
    pc = synth.uniform(seed, n, d) if seed % 2 == 0 else synth.grid_ties(seed, n, d, levels=5)
    np.testing.assert_array_equal(oracle.fps_vanilla(pc, k, s), ref.fps_sampling(pc, k, s))
    np.testing.assert_array_equal(oracle.kdline(pc, k, h, s), ref.bucket_fps_kdline_sampling(pc, k, h, s))
    starts = [int(x) for x in g.integers(0, n, size=min(4, k))]
    np.testing.assert_array_equal(oracle.fps_vanilla(pc, k, starts), ref.fps_sampling(pc, k, starts))


@pytest.mark.parametrize("seed", range(6))
def test_kdtree_is_vanilla_over_the_full_permutation(seed, oracle, ref):
    """bucket_fps_kdtree_sampling (src/_ext/KDTree.h:13-52) == exact FPS over the rows the full-depth kd build
    permuted, from POSITION start, ties to the highest position (the right child wins, KDNode.h:41-46)."""
    g = np.random.default_rng(300 + seed)
    n = int(g.integers(64, 5000))
    d = int(g.integers(1, 9))
    k = int(g.integers(1, n))
    s = int(g.integers(0, n))
    pc = (synth.uniform(seed, n, d), synth.grid_ties(seed, n, d, levels=4), synth.lidar(seed, n))[seed % 3] if d == 3 or seed % 3 < 2 \
        else synth.uniform(seed, n, d)
    np.testing.assert_array_equal(oracle.kdtree(pc, k, s), ref.bucket_fps_kdtree_sampling(pc, k, s))


def test_reference_error_codes(oracle):
    """src/wrapper.hpp:121-127: rc 1 = bad dim, rc 2 = bad start."""
    pc9 = synth.uniform(0, 100, 9)
    with pytest.raises(RuntimeError, match="error code 1"):
        oracle.kdline(pc9, 10, 2, 0)
    with pytest.raises(RuntimeError, match="error code 2"):
        oracle.kdline(synth.uniform(0, 100, 3), 10, 2, 100)


@pytest.mark.parametrize("gen,n,d,k,h,s", [("uniform", 3000, 3, 700, 5, 11), ("grid", 2000, 2, 600, 6, 0),
                                          ("grid", 1500, 6, 400, 4, 3), ("lidar", 5000, 3, 900, 7, 5)])
def test_lazy_equals_eager(gen, n, d, k, h, s, oracle):
    pc = {"uniform": lambda: synth.uniform(5, n, d), "grid": lambda: synth.grid_ties(5, n, d),
          "lidar": lambda: synth.lidar(5, n)}[gen]()
    np.testing.assert_array_equal(oracle.kdline(pc, k, h, s), oracle.kdline_eager(pc, k, h, s))


def test_kdline_is_exact_fps_over_permuted_array(oracle):
    """kd-line == vanilla recurrence started at perm[start] when no ties occur."""
    pc = synth.uniform(123, 2048, 3)
    out = oracle.kdline(pc, 256, 4, 17)
    perm, bounds, box = oracle.kdline_build(pc, 4)
    assert out[0] == perm[17]
    assert sorted(perm.tolist()) == list(range(2048))
    assert bounds[0] == 0 and bounds[-1] == 2048 and np.all(np.diff(bounds.astype(np.int64)) > 0)
    q = pc[perm]
    for L in range(len(bounds) - 1):  # tight boxes (src/_ext/KDTreeBase.h:181-207)
        seg = q[int(bounds[L]):int(bounds[L + 1])]
        np.testing.assert_array_equal(box[L, 0], seg.min(0))
        np.testing.assert_array_equal(box[L, 1], seg.max(0))
    np.testing.assert_array_equal(out, oracle.fps_vanilla(pc, 256, int(out[0])))


def test_certifier(oracle):
    pc = synth.uniform(9, 5000, 3)
    o = oracle.fps_vanilla(pc, 800, 4)
    assert oracle.certify_vanilla(pc, o) == (True, 0)
    bad = o.copy()
    bad[300] = (bad[300] + 1) % 5000
    ok, where = oracle.certify_vanilla(pc, bad)
    assert not ok and where == 300
    g = synth.grid_ties(1, 3000, 3)
    o = oracle.fps_vanilla(g, 500, [5, 1, 9])
    assert oracle.certify_vanilla(g, o, n_forced=3)[0]
    assert oracle.certify_vanilla(g, o, n_forced=3, n_threads=1)[0]
    o = oracle.kdline(g, 500, 6, 2)
    assert oracle.certify_kdline(g, o, 6, 2)[0]
    o[9] = o[8]
    assert not oracle.certify_kdline(g, o, 6, 2)[0]
