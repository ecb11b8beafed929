"""CPU suite: host-side logic and the drop-in boundary, no GPU compute.

 - libfps_b200.so loads and exports every symbol include/fps_b200.h declares (and nothing undeclared);
 - the python front-end mirrors the reference's argument / exception behaviour (src/fpsample/__init__.py,
   src/lib.cpp:52-109, 249-270, 522-579) for everything that is decided before device work starts;
 - without a GPU every compute entry fails LOUDLY (no CPU fallback), and the product never touches oracle/.
"""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import fpsample_b200 as fps
from fpsample_b200 import capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_GPU = capi.device_count() > 0


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "fps_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return set(re.findall(r"FPS_API\s+[\w\s\*]+?\b(\w+)\s*\(", txt))


def test_header_declares_expected_entry_points():
    syms = header_symbols()
    assert {"fps_b200_vanilla", "bucket_fps_kdline", "fps_b200_vanilla_batch", "fps_b200_kdline_batch"} <= syms
    assert syms == set(capi.EXPORTS), "capi.py binding table and include/fps_b200.h disagree"


def test_library_exports_every_declared_symbol():
    L = ctypes.CDLL(capi.LIB_PATH)
    for s in header_symbols():
        assert hasattr(L, s), f"{s} declared in include/fps_b200.h but not exported"
    out = subprocess.check_output(["nm", "-D", "--defined-only", capi.LIB_PATH], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    extra = {s for s in exported if not s.startswith("_")} - header_symbols()
    assert not extra, f"undeclared exports: {extra}"


def test_library_is_sm100a_only_and_torch_free():
    out = subprocess.check_output(["cuobjdump", "-lelf", capi.LIB_PATH], text=True)
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs
    ldd = subprocess.check_output(["ldd", capi.LIB_PATH], text=True)
    assert "torch" not in ldd and "python" not in ldd


def test_version_strings():
    assert b"sm_100a" in capi.lib().fps_b200_version()
    assert fps.__version__.startswith("1.0.2")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "fpsample_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|oracle/|liboracle", src, flags=re.M), fn


# ---- python front-end: reference behaviour decided before any device work ---------------------------------
PC = synth.uniform(0, 100, 3)


def test_assertions_like_reference():
    with pytest.raises(AssertionError):
        fps.fps_sampling(PC, 0, 0)                      # n_samples >= 1        (__init__.py:49)
    with pytest.raises(AssertionError):
        fps.fps_sampling(PC[0], 1, 0)                   # ndim == 2             (:50)
    with pytest.raises(AssertionError):
        fps.fps_sampling(PC, 101, 0)                    # n_pts >= n_samples    (:52)
    with pytest.raises(AssertionError):
        fps.fps_sampling(PC, 10, 100)                   # start range           (:53-56)
    with pytest.raises(AssertionError):
        fps.fps_sampling(PC, 2, [1, 2, 3])              # len(list) <= n_samples(:57-60)
    with pytest.raises(AssertionError):
        fps.bucket_fps_kdline_sampling(PC, 10, 0, 0)    # h >= 1                (:196)
    with pytest.raises(AssertionError):
        fps.bucket_fps_kdline_sampling(PC, 10, 7, 0)    # 2**h <= n_pts         (:197)
    with pytest.raises(TypeError):
        fps.bucket_fps_kdline_sampling(PC, 10, 3, [1, 2])  # list start dies in the range assert (:198-200)


def test_start_idx_types_like_reference():
    with pytest.raises(ValueError, match="start_idx should be None, int or list"):
        fps.fps_sampling(PC, 10, np.int64(3))           # numpy scalar is rejected (__init__.py:30-31)
    with pytest.raises(ValueError):
        fps.fps_sampling(PC, 10, 3.0)


def test_pybind_layer_errors_like_reference():
    m = fps._fpsample
    with pytest.raises(TypeError):
        m._fps_sampling(PC, 10, "0")                    # lib.cpp:260
    with pytest.raises(ValueError):
        m._fps_sampling(PC, 101, 0)                     # lib.cpp:76-81
    with pytest.raises(ValueError):
        m._fps_sampling(PC, 10, 100)                    # lib.cpp:83-107
    with pytest.raises(ValueError):
        m._fps_sampling(PC, 2, np.array([1, 2, 3], dtype=np.uint64))
    with pytest.raises(TypeError):
        m._bucket_fps_kdline_sampling(PC, 10, 3, "0")   # lib.cpp:534
    with pytest.raises(NotImplementedError):
        m._bucket_fps_kdline_sampling(PC, 10, 3, np.array([1], dtype=np.uint64))  # lib.cpp:537-540
    with pytest.raises(ValueError):
        m._bucket_fps_kdline_sampling(PC, 10, 3, 100)   # lib.cpp:545-547
    with pytest.raises(ValueError):
        m._bucket_fps_kdline_sampling(PC, 0, 3, 0)      # lib.cpp:548-553
    with pytest.raises(ValueError):
        m._bucket_fps_kdline_sampling(PC, 10, 0, 0)     # lib.cpp:554-557


def test_c_abi_return_codes_like_reference():
    """src/wrapper.hpp:121-127: rc 1 bad dim before rc 2 bad start; both decided before device work."""
    out = np.empty(4, dtype=np.uint64)
    pc9 = synth.uniform(0, 50, 9)
    L = capi.lib()
    assert L.bucket_fps_kdline(pc9.ctypes.data, 50, 9, 4, 99, 2, out.ctypes.data) == 1
    assert L.bucket_fps_kdline(PC.ctypes.data, 100, 3, 4, 100, 2, out.ctypes.data) == 2
    assert b"start_idx" in L.fps_b200_last_error()
    assert L.bucket_fps_kdline(PC.ctypes.data, 100, 3, 0, 0, 2, out.ctypes.data) == 3
    st = np.array([100], dtype=np.uint64)
    assert L.fps_b200_vanilla(PC.ctypes.data, 100, 3, 4, st.ctypes.data, 1, out.ctypes.data) == 2
    with pytest.raises(RuntimeError, match="failed with error code 1"):  # lib.cpp:574-576
        fps._fpsample._bucket_fps_kdline_sampling(pc9, 4, 2, 0)


def test_every_reference_entry_point_exists():
    """src/fpsample/__init__.py:209-217: all five public functions of the reference are there, with its signatures"""
    import inspect
    want = {"fps_sampling": ["pc", "n_samples", "start_idx"],
            "fps_npdu_sampling": ["pc", "n_samples", "w", "start_idx"],
            "fps_npdu_kdtree_sampling": ["pc", "n_samples", "w", "start_idx"],
            "bucket_fps_kdtree_sampling": ["pc", "n_samples", "start_idx"],
            "bucket_fps_kdline_sampling": ["pc", "n_samples", "h", "start_idx"]}
    for name, params in want.items():
        assert list(inspect.signature(getattr(fps, name)).parameters) == params, name
        assert name in fps.__all__
    with pytest.raises(AssertionError):
        fps.fps_npdu_kdtree_sampling(PC, 10**9)
    with pytest.warns(UserWarning, match="k is too large"):   # src/fpsample/__init__.py:136-138 (the cap is n_pts here, n_pts - 1 in the index-window variant)
        try:
            fps.fps_npdu_kdtree_sampling(PC, 10, w=10**9, start_idx=0)
        except RuntimeError:
            pass   # no GPU in the CPU suite: the call itself fails loudly after the front-end did its part


@pytest.mark.skipif(HAVE_GPU, reason="checks the no-GPU behaviour")
def test_no_cpu_fallback_without_gpu():
    assert capi.device_count() == 0
    for call in (lambda: fps.fps_sampling(PC, 10, 0), lambda: fps.bucket_fps_kdline_sampling(PC, 10, 3, 0),
                 lambda: fps.fps_sampling_batch(PC[None], 10), lambda: fps.bucket_fps_kdline_sampling_batch(PC[None], 10, 3),
                 lambda: fps.bucket_fps_kdtree_sampling(PC, 10, 0), lambda: fps.bucket_fps_kdtree_sampling_batch(PC[None], 10),
                 lambda: fps.fps_npdu_sampling(PC, 10, 8, 0)):
        with pytest.raises(RuntimeError, match="error code 4"):   # FPS_ERR_NO_DEVICE
            call()
    assert "no CPU fallback" in capi.lib().fps_b200_last_error().decode()


def test_cuda_array_interface_is_validated_on_the_host():
    """device arrays are taken as they are: float32, C-contiguous, right rank -- anything else is refused before any
    pointer reaches the library (no GPU needed for the checks)."""
    import fpsample_b200 as fps

    class Dev:
        def __init__(self, shape, typestr="<f4", strides=None):
            self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (0x7f0000000000, False),
                                             "version": 3, "strides": strides}

    assert fps._cuda_view(np.zeros((4, 3), np.float32), 2) is None           # host arrays take the usual path
    assert fps._cuda_view(Dev((2, 100, 3)), 3) == (0x7f0000000000, (2, 100, 3))
    assert fps._cuda_view(Dev((100, 3), strides=(12, 4)), 2) == (0x7f0000000000, (100, 3))
    with pytest.raises(TypeError):
        fps._cuda_view(Dev((100, 3), typestr="<f8"), 2)
    with pytest.raises(TypeError):
        fps._cuda_view(Dev((100, 3), strides=(4, 400)), 2)
    with pytest.raises(ValueError):
        fps._cuda_view(Dev((100, 3)), 3)
    with pytest.raises(AssertionError):
        fps.bucket_fps_kdline_sampling_batch(Dev((2, 100, 3)), 200, 3)       # n_samples > n_pts, before any device call


def test_streaming_sampler_wave_plan():
    """plan_kdline_stream (host arithmetic, no device): a batch is cut into FULL waves of narrow teams of warps (8 two-warp
    teams per SM) and a last partial wave on wide teams (4 four-warp teams per SM); wide records (> 4 dimensions) and more
    than 128 buckets need four-warp teams; the knobs override."""
    from fpsample_b200 import capi
    P = lambda B, d=3, h=7, sms=148: capi.describe_stream_plan(B, 100000, d, h, sms)
    assert P(512).startswith("512 clouds x WPC=4")                                   # one wave of four-warp teams (an 8-GPU shard of cfg 5)
    assert P(700).startswith("700 clouds x WPC=2")                                   # more than 592: one wave of two-warp teams
    assert P(1184).startswith("1184 clouds x WPC=2") and "+" not in P(1184)
    assert P(1200).startswith("1184 clouds x WPC=2") and "+ 16 clouds x WPC=4" in P(1200)
    assert P(4096).startswith("3552 clouds x WPC=2") and "+ 544 clouds x WPC=4" in P(4096)
    assert "WPC=2" not in P(4096, d=6) and "4096 clouds x WPC=4" in P(4096, d=6)      # 24-byte pending entries
    assert "WPC=2" not in P(4096, h=9) and "BPL=4" in P(4096, h=9)                   # 512 buckets: four per lane
    half = P(2048, sms=74)                                                          # waves scale with the SM count
    assert half.startswith("1776 clouds x WPC=2") and "+ 272 clouds x WPC=4" in half
    for B in (1, 3, 591, 593, 1183, 1185, 2367, 2369, 5000, 100000):                # every cloud is in exactly one launch
        parts = [int(x.split(" clouds")[0]) for x in P(B).split(" + ")]
        assert sum(parts) == B, P(B)
    with capi.tuning(stream_split=0):                                               # one team size per batch (the earlier plan)
        assert P(4096).startswith("4096 clouds x WPC=2") and P(512).startswith("512 clouds x WPC=4")
    with capi.tuning(stream_split=2):                                               # one-warp teams allowed: 16 per SM
        assert P(4096).startswith("2368 clouds x WPC=1")
    with capi.tuning(stream_warps=4):
        assert P(4096).startswith("4096 clouds x WPC=4")
    with pytest.raises(capi.FpsError):
        capi.describe_stream_plan(10, 1000, 9, 5)                                   # kd-line: at most 8 dimensions
