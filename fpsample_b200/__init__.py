"""fpsample_b200 -- B200-native farthest point sampling, drop-in for fpsample's FPS hot path.

Same Python surface as the reference front-end (src/fpsample/__init__.py:35-65, 174-206):

    fps_sampling(pc, n_samples, start_idx=None)                      -> uint64[n_samples]
    bucket_fps_kdline_sampling(pc, n_samples, h, start_idx=None)     -> uint64[n_samples]
    bucket_fps_kdtree_sampling(pc, n_samples, start_idx=None)        -> uint64[n_samples]   (:145-171)
    fps_npdu_sampling(pc, n_samples, w=None, start_idx=None)         -> uint64[n_samples]   (:66-103)

    fps_npdu_kdtree_sampling(pc, n_samples, w=None, start_idx=None)  -> uint64[n_samples]   (:106-142)

plus batched twins over [B, N, D] arrays (new):

    fps_sampling_batch(pcs, n_samples, start_idx=None, devices=None)               -> uint64[B, n_samples]
    bucket_fps_kdline_sampling_batch(pcs, n_samples, h, start_idx=None, devices=None)
    bucket_fps_kdtree_sampling_batch(pcs, n_samples, start_idx=None, devices=None)

Inputs may be numpy arrays or GPU-resident arrays (anything with __cuda_array_interface__: float32, C-contiguous).
Everything runs on the GPU through the C ABI in include/fps_b200.h; there is no CPU fallback and the
import fails loudly when the native extension has not been built (python build_native.py).
"""
from __future__ import annotations

import warnings
from typing import List, Optional, Sequence, Union

import numpy as np

try:
    from ._fpsample import (  # noqa: F401
        __version__,
        _bucket_fps_kdline_sampling,
        _bucket_fps_kdline_sampling_batch,
        _bucket_fps_kdtree_sampling,
        _bucket_fps_kdtree_sampling_batch,
        _batch_ptr,
        _device_count,
        _fps_npdu_kdtree_sampling,
        _fps_npdu_sampling,
        _fps_sampling,
        _fps_sampling_batch,
        _kernel_launches,
        _last_plan,
        _set_producer_stream,
    )
except ImportError as e:  # pragma: no cover - build problem, never a silent fallback
    raise ImportError(
        "fpsample_b200: the native CUDA extension is not built (run `python build_native.py`); "
        "there is no CPU fallback") from e


_VANILLA, _KDLINE, _KDTREE = 0, 1, 2


def _cuda_view(a, ndim: int):
    """(address, shape) of a GPU-resident array (anything with __cuda_array_interface__: torch, cupy, numba), else None.

    SURVEY.md 8(f) row 3: clouds that are already in HBM are sampled where they are.  No implicit casts or copies on
    the device: the buffer has to be C-contiguous float32.  Work that produces it is waited for ON THE DEVICE, through an
    event on the producer's stream -- the interface's `stream` entry, torch's current stream for torch tensors, else the
    legacy default stream -- never with a device-wide synchronisation.  The index array still comes back as a host array.
    """
    cai = getattr(a, "__cuda_array_interface__", None)
    if cai is None:
        return None
    shape = tuple(int(x) for x in cai["shape"])
    if len(shape) != ndim:
        raise ValueError(f"expected a {ndim}-D device array, got shape {shape}")
    if cai["typestr"] not in ("<f4", "=f4", "|f4"):
        raise TypeError("device arrays must be float32 (cast on the device before the call)")
    strides = cai.get("strides")
    if strides is not None:
        want, acc = [], 4
        for dim in reversed(shape):
            want.append(acc)
            acc *= dim
        if tuple(strides) != tuple(reversed(want)):
            raise TypeError("device arrays must be C-contiguous")
    _set_producer_stream(_producer_stream(a, cai))
    return int(cai["data"][0]), shape


def _producer_stream(a, cai) -> int:
    """cudaStream_t (as an int) whose queued work produces `a`; 0 = the legacy default stream"""
    st = cai.get("stream")
    if isinstance(st, int):
        return st          # 1 = legacy default, 2 = per-thread default (the CUDA handle values), else a stream handle
    if type(a).__module__.split(".")[0] == "torch":
        import torch
        return int(torch.cuda.current_stream(a.device).cuda_stream)
    return 0


def _device_single(algo: int, dv, n_samples: int, h: int, start_idx):
    ptr, (n_pts, d) = dv
    assert n_samples >= 1, "n_samples should be >= 1"
    assert n_pts >= n_samples, "n_pts should be >= n_samples"
    if start_idx is None:
        start_idx = int(np.random.randint(low=0, high=n_pts))
    if not isinstance(start_idx, int):
        raise TypeError("a device array takes start_idx None or int")
    assert 0 <= start_idx < n_pts, "start_idx should be None or 0 <= start_idx < n_pts"
    return _batch_ptr(algo, ptr, 1, n_pts, d, n_samples, h, start_idx, None)[0]


def get_start_idx(n_pts: int, start_idx: Optional[Union[int, List[int]]]) -> Union[int, np.ndarray]:
    """Reference: src/fpsample/__init__.py:19-32 (random start honours np.random.seed)."""
    if start_idx is None:
        start_idx = np.random.randint(low=0, high=n_pts)
    elif isinstance(start_idx, int):
        start_idx = start_idx
    elif isinstance(start_idx, list):
        start_idx = np.array(start_idx, dtype=np.uint64)
    else:
        raise ValueError("start_idx should be None, int or list")
    return start_idx


def fps_sampling(pc: np.ndarray, n_samples: int,
                 start_idx: Optional[Union[int, List[int]]] = None) -> np.ndarray:
    """Vanilla FPS (reference: src/fpsample/__init__.py:35-65 -> src/lib.cpp:188-246).

    Args:
        pc: point cloud of shape (n_pts, D); cast to float32.
        n_samples: number of samples.
        start_idx: None (random), int, or list[int] (forced first picks, all present in the result).
    Returns:
        uint64 indices of shape (n_samples,).
    """
    dv = _cuda_view(pc, 2)
    if dv is not None:
        return _device_single(_VANILLA, dv, n_samples, 0, start_idx)
    assert n_samples >= 1, "n_samples should be >= 1"
    assert pc.ndim == 2
    n_pts, _ = pc.shape
    assert n_pts >= n_samples, "n_pts should be >= n_samples"
    if isinstance(start_idx, int):
        assert start_idx is None or 0 <= start_idx < n_pts, "start_idx should be None or 0 <= start_idx < n_pts"
    if isinstance(start_idx, list):
        assert len(start_idx) <= n_samples, "len(start_idx) should be <= n_samples"
        for idx in start_idx:
            assert 0 <= idx < n_pts, "start_idx should be None or 0 <= start_idx < n_pts"
    pc = np.ascontiguousarray(pc, dtype=np.float32)
    start_idx = get_start_idx(n_pts, start_idx)
    return _fps_sampling(pc, n_samples, start_idx)


def bucket_fps_kdline_sampling(pc: np.ndarray, n_samples: int, h: int,
                               start_idx: Optional[Union[int, List[int]]] = None) -> np.ndarray:
    """QuickFPS with a kd-line of height h (reference: src/fpsample/__init__.py:174-206).

    As in the reference, start_idx addresses the POSITION in the array after the kd build permuted it
    (src/wrapper.hpp:54-55), so out[0] is generally not start_idx.
    """
    dv = _cuda_view(pc, 2)
    if dv is not None:
        assert h >= 1, "h should be >= 1"
        assert 2**h <= dv[1][0], "2**h should be <= n_pts"
        return _device_single(_KDLINE, dv, n_samples, h, start_idx)
    assert n_samples >= 1, "n_samples should be >= 1"
    assert pc.ndim == 2
    n_pts, _ = pc.shape
    assert n_pts >= n_samples, "n_pts should be >= n_samples"
    assert h >= 1, "h should be >= 1"
    assert 2**h <= n_pts, "2**h should be <= n_pts"
    assert start_idx is None or 0 <= start_idx < n_pts, "start_idx should be None or 0 <= start_idx < n_pts"
    if isinstance(start_idx, list):
        assert len(start_idx) <= n_samples, "len(start_idx) should be <= n_samples"
    pc = np.ascontiguousarray(pc, dtype=np.float32)
    start_idx = get_start_idx(n_pts, start_idx)
    return _bucket_fps_kdline_sampling(pc, n_samples, h, start_idx)


def bucket_fps_kdtree_sampling(pc: np.ndarray, n_samples: int,
                               start_idx: Optional[Union[int, List[int]]] = None) -> np.ndarray:
    """QuickFPS with the full kd tree (reference: src/fpsample/__init__.py:145-171).

    As in the reference, start_idx addresses the POSITION in the array after the (full-depth) kd build permuted
    it (src/wrapper.hpp:36-37), and distance ties go to the highest position (src/_ext/KDNode.h:41-46).
    """
    dv = _cuda_view(pc, 2)
    if dv is not None:
        return _device_single(_KDTREE, dv, n_samples, 0, start_idx)
    assert n_samples >= 1, "n_samples should be >= 1"
    assert pc.ndim == 2
    n_pts, _ = pc.shape
    assert n_pts >= n_samples, "n_pts should be >= n_samples"
    assert start_idx is None or 0 <= start_idx < n_pts, "start_idx should be None or 0 <= start_idx < n_pts"
    if isinstance(start_idx, list):
        assert len(start_idx) <= n_samples, "len(start_idx) should be <= n_samples"
    pc = np.ascontiguousarray(pc, dtype=np.float32)
    start_idx = get_start_idx(n_pts, start_idx)
    return _bucket_fps_kdtree_sampling(pc, n_samples, start_idx)


def _batch_start(start_idx, b: int, n_pts: int):
    if start_idx is None or isinstance(start_idx, int):
        if isinstance(start_idx, int):
            assert 0 <= start_idx < n_pts, "start_idx should be None or 0 <= start_idx < n_pts"
        return start_idx if start_idx is not None else 0
    arr = np.ascontiguousarray(start_idx, dtype=np.uint64)
    assert arr.shape == (b,), "start_idx should be None, int or a sequence of B ints"
    return arr


def fps_sampling_batch(pcs: np.ndarray, n_samples: int,
                       start_idx: Optional[Union[int, Sequence[int]]] = None,
                       devices: Optional[Sequence[int]] = None) -> np.ndarray:
    """Vanilla FPS over a batch [B, N, D]; row b equals fps_sampling(pcs[b], n_samples, start_idx[b]).

    start_idx None means 0 for every cloud (a batch is deterministic by default).  The batch is split
    into contiguous shards over `devices` (default: every visible B200); no inter-GPU traffic.
    """
    dv = _cuda_view(pcs, 3)
    if dv is not None:   # GPU-resident batch: sampled where it lives
        ptr, (b, n_pts, d) = dv
        assert n_samples >= 1, "n_samples should be >= 1"
        assert n_pts >= n_samples, "n_pts should be >= n_samples"
        return _batch_ptr(_VANILLA, ptr, b, n_pts, d, n_samples, 0, _batch_start(start_idx, b, n_pts),
                          None if devices is None else list(devices))
    assert n_samples >= 1, "n_samples should be >= 1"
    assert pcs.ndim == 3
    b, n_pts, _ = pcs.shape
    assert n_pts >= n_samples, "n_pts should be >= n_samples"
    pcs = np.ascontiguousarray(pcs, dtype=np.float32)
    return _fps_sampling_batch(pcs, n_samples, _batch_start(start_idx, b, n_pts),
                               None if devices is None else list(devices))


def bucket_fps_kdline_sampling_batch(pcs: np.ndarray, n_samples: int, h: int,
                                     start_idx: Optional[Union[int, Sequence[int]]] = None,
                                     devices: Optional[Sequence[int]] = None) -> np.ndarray:
    """QuickFPS kd-line over a batch [B, N, D]; row b equals bucket_fps_kdline_sampling(pcs[b], ...)."""
    dv = _cuda_view(pcs, 3)
    if dv is not None:   # GPU-resident batch: sampled where it lives
        ptr, (b, n_pts, d) = dv
        assert n_samples >= 1, "n_samples should be >= 1"
        assert n_pts >= n_samples, "n_pts should be >= n_samples"
        assert h >= 1, "h should be >= 1"
        assert 2**h <= n_pts, "2**h should be <= n_pts"
        return _batch_ptr(_KDLINE, ptr, b, n_pts, d, n_samples, h, _batch_start(start_idx, b, n_pts),
                          None if devices is None else list(devices))
    assert n_samples >= 1, "n_samples should be >= 1"
    assert pcs.ndim == 3
    b, n_pts, _ = pcs.shape
    assert n_pts >= n_samples, "n_pts should be >= n_samples"
    assert h >= 1, "h should be >= 1"
    assert 2**h <= n_pts, "2**h should be <= n_pts"
    pcs = np.ascontiguousarray(pcs, dtype=np.float32)
    return _bucket_fps_kdline_sampling_batch(pcs, n_samples, h, _batch_start(start_idx, b, n_pts),
                                             None if devices is None else list(devices))


def bucket_fps_kdtree_sampling_batch(pcs: np.ndarray, n_samples: int,
                                     start_idx: Optional[Union[int, Sequence[int]]] = None,
                                     devices: Optional[Sequence[int]] = None) -> np.ndarray:
    """QuickFPS full kd tree over a batch [B, N, D]; row b equals bucket_fps_kdtree_sampling(pcs[b], ...)."""
    dv = _cuda_view(pcs, 3)
    if dv is not None:   # GPU-resident batch: sampled where it lives
        ptr, (b, n_pts, d) = dv
        assert n_samples >= 1, "n_samples should be >= 1"
        assert n_pts >= n_samples, "n_pts should be >= n_samples"
        return _batch_ptr(_KDTREE, ptr, b, n_pts, d, n_samples, 0, _batch_start(start_idx, b, n_pts),
                          None if devices is None else list(devices))
    assert n_samples >= 1, "n_samples should be >= 1"
    assert pcs.ndim == 3
    b, n_pts, _ = pcs.shape
    assert n_pts >= n_samples, "n_pts should be >= n_samples"
    pcs = np.ascontiguousarray(pcs, dtype=np.float32)
    return _bucket_fps_kdtree_sampling_batch(pcs, n_samples, _batch_start(start_idx, b, n_pts),
                                             None if devices is None else list(devices))


def fps_npdu_sampling(pc: np.ndarray, n_samples: int, w: Optional[int] = None,
                      start_idx: Optional[Union[int, List[int]]] = None) -> np.ndarray:
    """FPS with the nearest-point-distance-updating heuristic over an index window (reference:
    src/fpsample/__init__.py:66-103 -> src/lib.cpp:272-366).  NOT exact FPS; needs dimensional locality, like the reference.

    w: window size of the local update, default n_pts / n_samples * 16 (capped to n_pts - 1 with the reference's warning).
    """
    assert n_samples >= 1, "n_samples should be >= 1"
    assert pc.ndim == 2
    n_pts, _ = pc.shape
    assert n_pts >= n_samples, "n_pts should be >= n_samples"
    assert start_idx is None or 0 <= start_idx < n_pts, "start_idx should be None or 0 <= start_idx < n_pts"
    if isinstance(start_idx, list):
        assert len(start_idx) <= n_samples, "len(start_idx) should be <= n_samples"
    pc = np.ascontiguousarray(pc, dtype=np.float32)
    w = w or int(n_pts / n_samples * 16)
    if w >= n_pts - 1:
        warnings.warn(f"k is too large, set to {n_pts - 1}")
        w = n_pts - 1
    start_idx = get_start_idx(n_pts, start_idx)
    return _fps_npdu_sampling(pc, n_samples, w, start_idx)


def fps_npdu_kdtree_sampling(pc: np.ndarray, n_samples: int, w: Optional[int] = None,
                             start_idx: Optional[Union[int, List[int]]] = None) -> np.ndarray:
    """FPS with the NPDU heuristic over the w NEAREST points of every pick instead of an index window (reference:
    src/fpsample/__init__.py:106-142 -> src/lib.cpp:369-465).  NOT exact FPS; needs no dimensional locality.

    w: number of neighbours updated per pick, default n_pts / n_samples * 16 (capped to n_pts with the reference's warning).
    The reference finds the neighbours with nanoflann; here the GPU selects them by brute force, which gives the same set --
    and the same indices -- unless several points share the w-th nearest distance exactly (see include/fps_b200.h).
    """
    assert n_samples >= 1, "n_samples should be >= 1"
    assert pc.ndim == 2
    n_pts, _ = pc.shape
    assert n_pts >= n_samples, "n_pts should be >= n_samples"
    assert start_idx is None or 0 <= start_idx < n_pts, "start_idx should be None or 0 <= start_idx < n_pts"
    if isinstance(start_idx, list):
        assert len(start_idx) <= n_samples, "len(start_idx) should be <= n_samples"
    pc = np.ascontiguousarray(pc, dtype=np.float32)
    w = w or int(n_pts / n_samples * 16)
    if w >= n_pts:
        warnings.warn(f"k is too large, set to {n_pts}")
        w = n_pts
    start_idx = get_start_idx(n_pts, start_idx)
    return _fps_npdu_kdtree_sampling(pc, n_samples, w, start_idx)


__all__ = [
    "__doc__",
    "__version__",
    "fps_sampling",
    "bucket_fps_kdline_sampling",
    "fps_sampling_batch",
    "bucket_fps_kdline_sampling_batch",
    "bucket_fps_kdtree_sampling_batch",
    "fps_npdu_sampling",
    "fps_npdu_kdtree_sampling",
    "bucket_fps_kdtree_sampling",
]
