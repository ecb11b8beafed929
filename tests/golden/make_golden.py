"""Generate tests/golden/golden.npz from the UNMODIFIED compiled reference (oracle/_ref/fpsample_ref, built
by oracle/build_ref.sh from /root/reference).  Run in the build container:  python tests/golden/make_golden.py

Stored per case: `<id>` = the reference's output indices (uint32) and `<id>__in` = sha256[:16] of the float32
input, so the tests notice if a seeded generator ever drifts.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from cases import CASES, input_sha, make_input  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    ref = O.load_reference()
    if ref is None:
        raise SystemExit("compiled reference missing: run `make -C oracle ref` where /root/reference exists")
    path = os.path.join(ROOT, "tests", "golden", "golden.npz")
    out = {}
    if "--all" not in sys.argv and os.path.exists(path):   # keep what is there, add the cases that are missing
        with np.load(path) as z:
            out = {k: z[k] for k in z.files}
    for cid, spec, call, p in CASES:
        if cid in out:
            continue
        pc = make_input(spec)
        t = time.time()
        if call == "vanilla":
            idx = ref.fps_sampling(pc, p["k"], p["start"])
        elif call == "npdu":
            idx = ref._fps_npdu_sampling(pc, p["k"], p["w"], p["start"])   # the binding itself: the python wrapper rewrites w
        elif call == "npdukd":
            idx = ref._fps_npdu_kdtree_sampling(pc, p["k"], p["w"], p["start"])
        elif call == "kdtree":
            idx = ref.bucket_fps_kdtree_sampling(pc, p["k"], p["start"])
        else:
            idx = ref.bucket_fps_kdline_sampling(pc, p["k"], p["h"], p["start"])
        assert idx.dtype == np.uint64 and idx.shape == (p["k"],)
        out[cid] = idx.astype(np.uint32)
        out[cid + "__in"] = np.array(input_sha(pc))
        print(f"{cid:28s} {time.time() - t:7.2f}s  first {idx[:4]}", flush=True)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
