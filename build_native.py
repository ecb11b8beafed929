"""In-tree build of the native pieces (no JIT cache: the .so files travel with the repo snapshot).

  libfps_b200.so            nvcc, sm_100a only: CUDA kernels + the C ABI (include/fps_b200.h)
  _fpsample.<abi>.so        g++: pybind11 module on top of the C ABI (host-side mirror of src/lib.cpp)

`python build_native.py [--force] [-v]` or `import build_native; build_native.build()`.
(Kept outside the package because the package refuses to import without its native extension.)
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

ROOT = os.path.dirname(os.path.abspath(__file__))
HERE = os.path.join(ROOT, "fpsample_b200")
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfps_b200.so")
EXT = os.path.join(HERE, "_fpsample" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false", *os.environ.get("FPS_NVCC_EXTRA", "").split(),  # bit-exact parity: no FMA contraction anywhere (the kernels also use *_rn)
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]
CU = ["vanilla.cu", "kdtree.cu", "kdline.cu", "kdsmall.cu", "seqsum.cu", "npdu.cu", "kdline_async.cu", "kdline_warp.cu", "kdline_stream.cu", "kdline_grid.cu", "kdbuild.cu", "comm.cu", "floors.cu", "capi.cu"]
HDR = ["common.cuh", "kdcommon.cuh", "seqsum.cuh", "engine.h", os.path.join(ROOT, "include", "fps_b200.h")]


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> None:
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    hdrs = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HDR]
    objs, jobs = [], []
    for cu in CU:
        src = os.path.join(CSRC, cu)
        obj = os.path.join(objdir, cu.replace(".cu", ".o"))
        objs.append(obj)
        if force or _newer(obj, [src] + hdrs):
            jobs.append([NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj])
    if jobs:   # translation units are independent: compile them side by side
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            list(ex.map(subprocess.check_call, jobs))
    if force or _newer(LIB, objs):
        # default (static) cudart: the library does not depend on which libcudart the host process loaded
        subprocess.check_call([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs +
                              ["-Xcompiler", "-fPIC", "-ldl"])
    src = os.path.join(CSRC, "pymodule.cpp")
    if force or _newer(EXT, [src, LIB] + hdrs):
        import pybind11
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden",
               "-I" + sysconfig.get_paths()["include"], "-I" + pybind11.get_include(),
               src, "-o", EXT, "-L" + HERE, "-lfps_b200", "-Wl,-rpath,$ORIGIN"]
        subprocess.check_call(cmd)


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", LIB, "and", EXT)
