"""ctypes front-end of the C oracle (oracle/fps_oracle.c) and loader of the compiled reference.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module; fpsample_b200 never does (tests/test_host.py
greps for it).  Parity pinning: see the header of fps_oracle.c.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C oracle with gcc (seconds).  Returns the path of the shared object."""
    src = os.path.join(_HERE, "fps_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_LIB), exist_ok=True)
        subprocess.check_call(
            ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
             src, "-o", _LIB, "-lm", "-lpthread"])
    return _LIB


def _load():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB)
        sz, fp, szp = ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p
        L.oracle_fps_vanilla.argtypes = [fp, sz, sz, sz, szp, sz, szp]
        L.oracle_kdline_build.argtypes = [fp, sz, sz, sz, szp, szp, fp, szp]
        L.oracle_kdline_sample.argtypes = [fp, sz, sz, sz, sz, sz, szp, ctypes.c_void_p]
        L.oracle_kdline_sample_eager.argtypes = [fp, sz, sz, sz, sz, sz, szp]
        L.oracle_certify_fps.argtypes = [fp, sz, sz, sz, szp, sz, ctypes.c_int, ctypes.c_int, szp]
        L.oracle_fps_npdu.argtypes = [fp, sz, sz, sz, sz, sz, szp]
        L.oracle_fps_npdu.restype = ctypes.c_int
        L.oracle_fps_npdu_kdtree.argtypes = [fp, sz, sz, sz, sz, sz, szp]
        L.oracle_fps_npdu_kdtree.restype = ctypes.c_int
        for f in (L.oracle_fps_vanilla, L.oracle_kdline_build, L.oracle_kdline_sample,
                  L.oracle_kdline_sample_eager, L.oracle_certify_fps):
            f.restype = ctypes.c_int
        _lib = L
    return _lib


def _f32(pc):
    pc = np.ascontiguousarray(pc, dtype=np.float32)
    assert pc.ndim == 2
    return pc


def _check(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed with error code {rc}")


def fps_vanilla(pc, k, start=0):
    """Oracle twin of fpsample.fps_sampling (start: int or list of ints)."""
    pc = _f32(pc)
    starts = np.atleast_1d(np.asarray(start, dtype=np.uint64)).copy()
    out = np.empty(k, dtype=np.uint64)
    rc = _load().oracle_fps_vanilla(pc.ctypes.data, pc.shape[0], pc.shape[1], k, starts.ctypes.data,
                                    starts.size, out.ctypes.data)
    _check(rc, "oracle_fps_vanilla")
    return out


def fps_npdu(pc, n_samples, w, start=0):
    """Oracle twin of fpsample._fpsample._fps_npdu_sampling (index-window heuristic, src/lib.cpp:272-340)."""
    pc = _f32(pc)
    out = np.empty(n_samples, dtype=np.uint64)
    rc = _load().oracle_fps_npdu(pc.ctypes.data, pc.shape[0], pc.shape[1], n_samples, w, start, out.ctypes.data)
    _check(rc, "oracle_fps_npdu")
    return out


def kdline_build(pc, h):
    """-> (perm[N] uint64, leaf_bounds[n_leaves+1] uint64, leaf_box[n_leaves,2,D] float32)."""
    pc = _f32(pc)
    n, d = pc.shape
    perm = np.empty(n, dtype=np.uint64)
    cap = min(1 << min(h, 40), n)
    bounds = np.empty(cap + 1, dtype=np.uint64)
    box = np.empty((cap, 2, d), dtype=np.float32)
    nl = ctypes.c_size_t(0)
    rc = _load().oracle_kdline_build(pc.ctypes.data, n, d, h, perm.ctypes.data, bounds.ctypes.data,
                                     box.ctypes.data, ctypes.addressof(nl))
    _check(rc, "oracle_kdline_build")
    return perm, bounds[: nl.value + 1].copy(), box[: nl.value].copy()


def fps_npdu_kdtree(pc, n_samples, k, start=0):
    """Oracle twin of fpsample._fpsample._fps_npdu_kdtree_sampling (k-nearest-neighbour heuristic, src/lib.cpp:369-465)."""
    pc = _f32(pc)
    out = np.empty(n_samples, dtype=np.uint64)
    rc = _load().oracle_fps_npdu_kdtree(pc.ctypes.data, pc.shape[0], pc.shape[1], n_samples, k, start, out.ctypes.data)
    _check(rc, "oracle_fps_npdu_kdtree")
    return out


def kdline(pc, k, h, start=0, return_stats=False):
    """Oracle twin of fpsample.bucket_fps_kdline_sampling (lazy bucket form)."""
    pc = _f32(pc)
    out = np.empty(k, dtype=np.uint64)
    stats = np.zeros(5, dtype=np.uint64)
    rc = _load().oracle_kdline_sample(pc.ctypes.data, pc.shape[0], pc.shape[1], k, start, h,
                                      out.ctypes.data, stats.ctypes.data)
    _check(rc, "oracle_kdline_sample")
    if return_stats:
        return out, dict(zip(("point_updates", "bucket_tests", "flushes", "deferred", "dropped"),
                             (int(x) for x in stats)))
    return out


def kdline_eager(pc, k, h, start=0):
    """Exact FPS over the permuted array, O(N*K)."""
    pc = _f32(pc)
    out = np.empty(k, dtype=np.uint64)
    rc = _load().oracle_kdline_sample_eager(pc.ctypes.data, pc.shape[0], pc.shape[1], k, start, h,
                                            out.ctypes.data)
    _check(rc, "oracle_kdline_sample_eager")
    return out


def kdtree(pc, k, start=0):
    """Oracle twin of fpsample.bucket_fps_kdtree_sampling (src/_ext/KDTree.h:13-52, src/wrapper.hpp:29-43): the build
    permutes the array down to single points (KDTree.h:27, same split rules as the kd-line, KDTreeBase.h:84-207),
    sampling starts at POSITION start, the right child wins distance ties (KDNode.h:41-46) = highest position: exact
    FPS over the fully permuted rows with vanilla's tie rule.  Pinned against the compiled reference's outputs
    (tests/golden: *_kdtree cases) and, where oracle/_ref exists, live (tests/test_oracle.py)."""
    pc = _f32(pc)
    perm, _, _ = kdline_build(pc, 1 << 20)   # depth limit far beyond any tree: leaves are single points
    q = np.ascontiguousarray(pc[perm.astype(np.int64)])
    return perm[fps_vanilla(q, k, start).astype(np.int64)].astype(np.uint64)


def certify_vanilla(pc, picks, n_forced=1, n_threads=None):
    """True iff `picks` is exactly what fps_sampling yields from picks[:n_forced] (all host cores)."""
    pc = _f32(pc)
    picks = np.ascontiguousarray(picks, dtype=np.uint64)
    bad = ctypes.c_size_t(0)
    rc = _load().oracle_certify_fps(pc.ctypes.data, pc.shape[0], pc.shape[1], picks.size, picks.ctypes.data,
                                    n_forced, 0, n_threads or os.cpu_count() or 1, ctypes.addressof(bad))
    if rc not in (0, 1):
        raise RuntimeError(f"oracle_certify_fps failed with error code {rc}")
    return rc == 0, int(bad.value)


def certify_kdline(pc, picks, h, start=0, n_threads=None):
    """True iff `picks` is what bucket_fps_kdline_sampling(pc, len(picks), h, start) yields: builds the
    oracle permutation, maps the picks to positions and certifies exact FPS over the permuted rows."""
    pc = _f32(pc)
    perm, _, _ = kdline_build(pc, h)
    inv = np.empty(perm.size, dtype=np.uint64)
    inv[perm] = np.arange(perm.size, dtype=np.uint64)
    picks = np.ascontiguousarray(picks, dtype=np.uint64)
    if picks.size and (picks.max() >= perm.size or inv[picks[0]] != start):
        return False, 0
    pos = np.ascontiguousarray(inv[picks])
    q = np.ascontiguousarray(pc[perm])
    bad = ctypes.c_size_t(0)
    rc = _load().oracle_certify_fps(q.ctypes.data, q.shape[0], q.shape[1], pos.size, pos.ctypes.data, 1, 1,
                                    n_threads or os.cpu_count() or 1, ctypes.addressof(bad))
    if rc not in (0, 1):
        raise RuntimeError(f"oracle_certify_fps failed with error code {rc}")
    return rc == 0, int(bad.value)


def load_reference():
    """Import the unmodified compiled reference (oracle/_ref/fpsample_ref) or return None."""
    ref_dir = os.path.join(_HERE, "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "fpsample_ref")):
        return None
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    try:
        import fpsample_ref  # type: ignore
        return fpsample_ref
    except Exception:  # pragma: no cover - e.g. ABI mismatch on another interpreter
        return None
