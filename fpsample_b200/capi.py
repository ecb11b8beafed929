"""ctypes binding of the C ABI (include/fps_b200.h) -- what a non-pybind host (or a framework holding
device pointers) would bind.  Used by the parity tests and bench.py; no compute happens in Python.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FPS_B200_LIB") or os.path.join(_HERE, "libfps_b200.so")   # FPS_B200_LIB: an experimental build (scripts/build_variant.sh)

ALGO_VANILLA, ALGO_KDLINE, ALGO_KDTREE = 0, 1, 2

EXPORTS = {
    # name: (restype, argtypes)
    "fps_b200_vanilla": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_size_t] * 3 + [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "bucket_fps_kdline": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_size_t] * 5 + [ctypes.c_void_p]),
    "bucket_fps_kdtree": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_size_t] * 4 + [ctypes.c_void_p]),
    "fps_b200_kdtree_batch": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_size_t] * 4 + [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "fps_b200_kdtree_batch_dev": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_size_t] * 4 + [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "fps_b200_vanilla_batch": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_size_t] * 4 + [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "fps_b200_kdline_batch": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_size_t] * 4 + [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "fps_b200_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int] + [ctypes.c_size_t] * 5),
    "fps_b200_vanilla_batch_dev": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_size_t] * 4 + [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "fps_b200_kdline_batch_dev": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_size_t] * 4 + [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "fps_b200_kdline_build_dev": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_size_t] * 4 + [ctypes.c_void_p] * 4 + [ctypes.c_size_t, ctypes.c_void_p]),
    "fps_b200_npdu": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_size_t] * 5 + [ctypes.c_void_p]),
    "fps_b200_npdu_batch": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_size_t] * 5 + [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "fps_b200_npdu_kdtree": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_size_t] * 5 + [ctypes.c_void_p]),
    "fps_b200_npdu_kdtree_batch": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_size_t] * 5 + [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "fps_b200_seqsum_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "fps_b200_device_count": (ctypes.c_int, []),
    "fps_b200_version": (ctypes.c_char_p, []),
    "fps_b200_last_error": (ctypes.c_char_p, []),
    "fps_b200_last_plan": (ctypes.c_char_p, []),
    "fps_b200_kernel_launches": (ctypes.c_uint64, []),
    "fps_b200_debug_counters": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p]),
    "fps_b200_comm_unique_id": (ctypes.c_int, [ctypes.c_void_p]),
    "fps_b200_comm_init": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]),
    "fps_b200_comm_init_local": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "fps_b200_comm_destroy": (None, []),
    "fps_b200_comm_ranks": (ctypes.c_int, []),
    "fps_b200_nccl_version": (ctypes.c_int, []),
    "fps_b200_nccl_library": (ctypes.c_int, [ctypes.c_char_p]),
    "fps_b200_kdline_batch_sharded": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_size_t] * 4 + [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "fps_b200_vanilla_batch_sharded": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_size_t] * 4 + [ctypes.c_void_p, ctypes.c_void_p]),
    "fps_b200_gather_indices": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_size_t] * 3 + [ctypes.c_void_p]),
    "fps_b200_set_tuning": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_long]),
    "fps_b200_set_producer_stream": (None, [ctypes.c_void_p]),
    "fps_b200_phase_timing": (None, [ctypes.c_int]),
    "fps_b200_last_phase_ms": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "fps_b200_sync_floor": (ctypes.c_int, [ctypes.c_int] * 4 + [ctypes.c_void_p]),
    "fps_b200_describe_stream_plan": (ctypes.c_int, [ctypes.c_size_t] * 4 + [ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t]),
    "fps_b200_host_alloc": (ctypes.c_void_p, [ctypes.c_size_t]),
    "fps_b200_host_free": (None, [ctypes.c_void_p]),
}

_lib = None


def lib():
    """Load libfps_b200.so (fails loudly if it was not built: there is no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python build_native.py` (no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
        p = _bundled_nccl()
        if p:   # the copy PyTorch would load if it is imported after the first comm call (include/fps_b200.h)
            L.fps_b200_nccl_library(p.encode())
    return _lib


def _bundled_nccl():
    """Path of the pip-installed nvidia-nccl wheel's libnccl.so.2, if there is one (nothing is imported)."""
    import importlib.util
    try:
        spec = importlib.util.find_spec("nvidia.nccl")
    except (ImportError, ValueError):
        return None
    for d in (spec.submodule_search_locations or []) if spec else []:
        p = os.path.join(d, "lib", "libnccl.so.2")
        if os.path.exists(p):
            return p
    return None


class FpsError(RuntimeError):
    def __init__(self, fn, rc):
        self.rc = rc
        super().__init__(f"{fn} failed with error code {rc}: {lib().fps_b200_last_error().decode()}")


def _check(fn, rc):
    if rc != 0:
        raise FpsError(fn, rc)


def last_plan() -> str:
    return lib().fps_b200_last_plan().decode()


def kernel_launches() -> int:
    return int(lib().fps_b200_kernel_launches())


DBG_ASYNC, DBG_WARP, DBG_BUILD, DBG_GRID, DBG_STREAM = 0, 1, 2, 3, 4


def debug_counters(which: int = DBG_ASYNC):
    """16 raw counters of the last launch of one sampler family (diagnostics / bench.py's executed-work roofline)."""
    out = np.zeros(16, dtype=np.uint64)
    _check("fps_b200_debug_counters", lib().fps_b200_debug_counters(which, out.ctypes.data))
    if which != DBG_ASYNC:
        return out
    names = ("iterations", "picks", "stalled_iterations", "cyc_poll_warptop", "cyc_wait_warps", "cyc_blocktop",
             "cyc_tests")
    return {k: int(v) for k, v in zip(names, out)}


def set_tuning(name: str, value: int = -1) -> None:
    """Planner override (value -1 = the planner's own choice); the environment is only read once, at first use."""
    _check("fps_b200_set_tuning", lib().fps_b200_set_tuning(name.upper().encode(), int(value)))


class tuning:
    """`with capi.tuning(group=0, warp_global_minb=1): ...` -- overrides for the block, planner defaults afterwards."""

    def __init__(self, **knobs):
        self.knobs = knobs

    def __enter__(self):
        for k, v in self.knobs.items():
            set_tuning(k, v)
        return self

    def __exit__(self, *exc):
        for k in self.knobs:
            set_tuning(k, -1)
        return False


def set_producer_stream(stream: int) -> None:
    """The calling thread's next host-pointer-entry call with a DEVICE input waits for `stream` (a cudaStream_t) on the device."""
    lib().fps_b200_set_producer_stream(stream or None)


def phase_timing(enable: bool) -> None:
    lib().fps_b200_phase_timing(1 if enable else 0)


def last_phase_ms():
    """(build_ms, sample_ms) of this thread's last *_dev call made with phase timing enabled."""
    b, s = ctypes.c_float(0), ctypes.c_float(0)
    _check("fps_b200_last_phase_ms", lib().fps_b200_last_phase_ms(ctypes.byref(b), ctypes.byref(s)))
    return float(b.value), float(s.value)


FLOOR_WARP, FLOOR_CLUSTER, FLOOR_GRID = 0, 1, 2


def sync_floor(kind: int, ctas: int = 1, words: int = 0, group: int = 0, rounds: int = 2000) -> float:
    """nanoseconds per empty round of a sampler's synchronisation structure (csrc/floors.cu)"""
    ns = ctypes.c_float(0)
    _check("fps_b200_sync_floor", lib().fps_b200_sync_floor(kind, ctas, (group << 16) | words, rounds, ctypes.byref(ns)))
    return float(ns.value)


def describe_stream_plan(n_clouds: int, n: int, dim: int, height: int, n_sms: int = 148) -> str:
    """how the streaming sampler cuts a batch into launches (host arithmetic only: works without a GPU)"""
    buf = ctypes.create_string_buffer(256)
    _check("fps_b200_describe_stream_plan", lib().fps_b200_describe_stream_plan(n_clouds, n, dim, height, n_sms, buf, len(buf)))
    return buf.value.decode()


def device_count() -> int:
    return int(lib().fps_b200_device_count())


def _f32(a, ndim):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == ndim
    return a


# ---- host-pointer entries -------------------------------------------------------------------------------
def vanilla(pc, k, starts=0):
    pc = _f32(pc, 2)
    st = np.atleast_1d(np.asarray(starts, dtype=np.uint64)).copy()
    out = np.empty(k, dtype=np.uint64)
    _check("fps_b200_vanilla", lib().fps_b200_vanilla(pc.ctypes.data, pc.shape[0], pc.shape[1], k, st.ctypes.data,
                                                      st.size, out.ctypes.data))
    return out


def kdline(pc, k, h, start=0):
    pc = _f32(pc, 2)
    out = np.empty(k, dtype=np.uint64)
    _check("bucket_fps_kdline", lib().bucket_fps_kdline(pc.ctypes.data, pc.shape[0], pc.shape[1], k, start, h,
                                                        out.ctypes.data))
    return out


def kdtree(pc, k, start=0):
    pc = _f32(pc, 2)
    out = np.empty(k, dtype=np.uint64)
    _check("bucket_fps_kdtree", lib().bucket_fps_kdtree(pc.ctypes.data, pc.shape[0], pc.shape[1], k, start, out.ctypes.data))
    return out


def npdu(pc, k, w, start=0):
    pc = _f32(pc, 2)
    out = np.empty(k, dtype=np.uint64)
    _check("fps_b200_npdu", lib().fps_b200_npdu(pc.ctypes.data, pc.shape[0], pc.shape[1], k, w, start, out.ctypes.data))
    return out


def npdu_kdtree(pc, k, w, start=0):
    pc = _f32(pc, 2)
    out = np.empty(k, dtype=np.uint64)
    _check("fps_b200_npdu_kdtree", lib().fps_b200_npdu_kdtree(pc.ctypes.data, pc.shape[0], pc.shape[1], k, w, start, out.ctypes.data))
    return out


def npdu_kdtree_batch(pcs, k, w, start=None, devices=None):
    pcs = _f32(pcs, 3)
    b, n, d = pcs.shape
    st = _starts(start, b)
    dv = None if devices is None else np.asarray(devices, dtype=np.int32)
    out = np.empty((b, k), dtype=np.uint64)
    _check("fps_b200_npdu_kdtree_batch", lib().fps_b200_npdu_kdtree_batch(
        pcs.ctypes.data, b, n, d, k, w, None if st is None else st.ctypes.data, out.ctypes.data,
        None if dv is None else dv.ctypes.data, 0 if dv is None else dv.size))
    return out


def npdu_batch(pcs, k, w, start=None, devices=None):
    pcs = _f32(pcs, 3)
    b, n, d = pcs.shape
    st = _starts(start, b)
    dv = None if devices is None else np.asarray(devices, dtype=np.int32)
    out = np.empty((b, k), dtype=np.uint64)
    _check("fps_b200_npdu_batch", lib().fps_b200_npdu_batch(
        pcs.ctypes.data, b, n, d, k, w, None if st is None else st.ctypes.data, out.ctypes.data,
        None if dv is None else dv.ctypes.data, 0 if dv is None else dv.size))
    return out


def _starts(start, b):
    if start is None:
        return None
    st = np.ascontiguousarray(np.broadcast_to(np.asarray(start, dtype=np.uint64), (b,)))
    return st


def vanilla_batch(pcs, k, start=None, devices=None):
    pcs = _f32(pcs, 3)
    b, n, d = pcs.shape
    st = _starts(start, b)
    dv = None if devices is None else np.asarray(devices, dtype=np.int32)
    out = np.empty((b, k), dtype=np.uint64)
    _check("fps_b200_vanilla_batch", lib().fps_b200_vanilla_batch(
        pcs.ctypes.data, b, n, d, k, None if st is None else st.ctypes.data, out.ctypes.data,
        None if dv is None else dv.ctypes.data, 0 if dv is None else dv.size))
    return out


def kdline_batch(pcs, k, h, start=None, devices=None):
    pcs = _f32(pcs, 3)
    b, n, d = pcs.shape
    st = _starts(start, b)
    dv = None if devices is None else np.asarray(devices, dtype=np.int32)
    out = np.empty((b, k), dtype=np.uint64)
    _check("fps_b200_kdline_batch", lib().fps_b200_kdline_batch(
        pcs.ctypes.data, b, n, d, k, None if st is None else st.ctypes.data, h, out.ctypes.data,
        None if dv is None else dv.ctypes.data, 0 if dv is None else dv.size))
    return out


def kdtree_batch(pcs, k, start=None, devices=None):
    pcs = _f32(pcs, 3)
    b, n, d = pcs.shape
    st = _starts(start, b)
    dv = None if devices is None else np.asarray(devices, dtype=np.int32)
    out = np.empty((b, k), dtype=np.uint64)
    _check("fps_b200_kdtree_batch", lib().fps_b200_kdtree_batch(
        pcs.ctypes.data, b, n, d, k, None if st is None else st.ctypes.data, out.ctypes.data,
        None if dv is None else dv.ctypes.data, 0 if dv is None else dv.size))
    return out


# ---- multi-GPU: NCCL gather of the index arrays to rank 0 (include/fps_b200.h, csrc/comm.cu) ---------------------------------
def comm_unique_id() -> bytes:
    buf = ctypes.create_string_buffer(128)
    _check("fps_b200_comm_unique_id", lib().fps_b200_comm_unique_id(buf))
    return buf.raw


def comm_init(uid: bytes, n_ranks: int, rank: int) -> None:
    assert len(uid) == 128
    _check("fps_b200_comm_init", lib().fps_b200_comm_init(ctypes.create_string_buffer(uid, 128), n_ranks, rank))


def comm_init_local(devices) -> None:
    dv = np.asarray(devices, dtype=np.int32)
    _check("fps_b200_comm_init_local", lib().fps_b200_comm_init_local(dv.ctypes.data, dv.size))


def comm_destroy() -> None:
    lib().fps_b200_comm_destroy()


def comm_ranks() -> int:
    return int(lib().fps_b200_comm_ranks())


def nccl_version() -> int:
    return int(lib().fps_b200_nccl_version())


_root_cache = {}   # shape -> up to 3 page-locked result buffers


def _root_out(shape):
    """rank 0's result buffer: page-locked when the driver hands it out (the copy from the GPU then runs at PCIe speed).
    Pinning a quarter of a gigabyte costs ~0.1 s, so a few buffers per shape are kept and one is handed out again once the
    caller has dropped the array it got (a surviving SLICE of it does not count: keep the array itself)."""
    import sys
    pool = _root_cache.setdefault(tuple(shape), [])
    for arr in pool:
        if sys.getrefcount(arr) <= 3:   # the pool, `arr`, getrefcount's argument: nobody else holds it
            return arr
    try:
        arr = pinned_empty(shape, np.uint64)
    except MemoryError:
        return np.empty(shape, dtype=np.uint64)
    if len(pool) < 3:
        pool.append(arr)
    return arr


def _sharded(fn, name, pcs, n_clouds, k, start, is_root, *extra):
    """pcs: this rank's shard (one process per GPU) or the whole batch (comm_init_local); numpy array or an int device address
    with `shape`.  -> [n_clouds, k] uint64 where rank 0 lives (pinned), else None."""
    if isinstance(pcs, tuple):
        ptr, (b, n, d) = pcs
        keep = None
    else:
        keep = pcs = _f32(pcs, 3)
        ptr, (b, n, d) = pcs.ctypes.data, pcs.shape
    st = _starts(start, b)
    out = _root_out((n_clouds, k)) if is_root else None
    args = [ptr, n_clouds, n, d, k, None if st is None else st.ctypes.data] + list(extra) + [None if out is None else out.ctypes.data]
    _check(name, fn(*args))
    del keep
    return out


def kdline_batch_sharded(pcs, n_clouds, k, h, start=None, is_root=True):
    return _sharded(lib().fps_b200_kdline_batch_sharded, "fps_b200_kdline_batch_sharded", pcs, n_clouds, k, start, is_root, h)


def vanilla_batch_sharded(pcs, n_clouds, k, start=None, is_root=True):
    return _sharded(lib().fps_b200_vanilla_batch_sharded, "fps_b200_vanilla_batch_sharded", pcs, n_clouds, k, start, is_root)


def gather_indices(local, n_clouds, is_root=True):
    local = np.ascontiguousarray(local, dtype=np.uint64)
    nb, k = local.shape
    out = _root_out((n_clouds, k)) if is_root else None
    _check("fps_b200_gather_indices", lib().fps_b200_gather_indices(local.ctypes.data, nb, k, n_clouds, None if out is None else out.ctypes.data))
    return out


# ---- device-pointer entries (raw integer addresses, e.g. torch.Tensor.data_ptr()) -------------------------
def workspace_bytes(algo, b, n, d, k, h=0) -> int:
    return int(lib().fps_b200_workspace_bytes(algo, b, n, d, k, h))


def vanilla_batch_dev(d_pts, b, n, d, k, d_start, d_out, d_ws, ws_bytes, stream=0):
    _check("fps_b200_vanilla_batch_dev", lib().fps_b200_vanilla_batch_dev(
        d_pts, b, n, d, k, d_start or None, d_out, d_ws or None, ws_bytes, stream or None))


def kdline_batch_dev(d_pts, b, n, d, k, d_start, h, d_out, d_ws, ws_bytes, stream=0):
    _check("fps_b200_kdline_batch_dev", lib().fps_b200_kdline_batch_dev(
        d_pts, b, n, d, k, d_start or None, h, d_out, d_ws or None, ws_bytes, stream or None))


def kdtree_batch_dev(d_pts, b, n, d, k, d_start, d_out, d_ws, ws_bytes, stream=0):
    _check("fps_b200_kdtree_batch_dev", lib().fps_b200_kdtree_batch_dev(
        d_pts, b, n, d, k, d_start or None, d_out, d_ws or None, ws_bytes, stream or None))


def kdline_build_dev(d_pts, b, n, d, h, d_perm, d_leaf_lo, d_leaf_box, d_ws, ws_bytes, stream=0):
    _check("fps_b200_kdline_build_dev", lib().fps_b200_kdline_build_dev(
        d_pts, b, n, d, h, d_perm, d_leaf_lo or None, d_leaf_box or None, d_ws or None, ws_bytes, stream or None))


def seqsum_dev(d_values, n, d_sum, d_fast_tiles=0, tile=512, stream=0):
    _check("fps_b200_seqsum_dev", lib().fps_b200_seqsum_dev(d_values, n, d_sum, d_fast_tiles or None, tile, stream or None))


def pinned_empty(shape, dtype):
    """numpy array backed by page-locked memory from the library (freed when the array is collected)."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    p = lib().fps_b200_host_alloc(max(nbytes, 1))
    if not p:
        raise MemoryError("fps_b200_host_alloc failed")
    buf = (ctypes.c_byte * max(nbytes, 1)).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    import weakref
    weakref.finalize(buf, lib().fps_b200_host_free, p)
    return arr
