"""Small calls through every sampler family for compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool racecheck python scripts/san_small.py
Each result is compared with the oracle, so a race that changes an index shows up even where the tool stays silent."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fpsample_b200 import capi, synth
from oracle import oracle as O

def check(name, got, want):
    ok = np.array_equal(np.asarray(got), np.asarray(want))
    print(f"{name:34s} {'OK ' if ok else 'MISMATCH'} | {capi.last_plan()[:110]}", flush=True)
    assert ok, name

pcs = synth.uniform_batch(5, 6, 2048, 3)
check("on-chip warp kernel", capi.kdline_batch(pcs, 64, 5, None, devices=[0]), np.stack([O.kdline(p, 64, 5, 0) for p in pcs]))
lat = np.stack([synth.grid_ties(40 + b, 1500, 2) for b in range(4)])
check("on-chip, tie lattice", capi.kdline_batch(lat, 48, 6, None, devices=[0]), np.stack([O.kdline(p, 48, 6, 0) for p in lat]))
big = synth.uniform_batch(9, 2, 16384, 3)
check("grouped grid kernel", capi.kdline_batch(big, 64, 7, None, devices=[0]), np.stack([O.kdline(p, 64, 7, 0) for p in big]))
for wpc in (1, 2, 4):
    with capi.tuning(group=0, warp_global_minb=1, stream_warps=wpc):
        st = synth.uniform_batch(11, 3, 20000, 3)
        check(f"streaming kernel, {wpc} warp(s)/cloud", capi.kdline_batch(st, 96, 7, None, devices=[0]), np.stack([O.kdline(p, 96, 7, 0) for p in st]))
with capi.tuning(group=0, warp_global_minb=1):
    s6 = synth.uniform_batch(12, 2, 9000, 6)
    check("streaming kernel, 6-D", capi.kdline_batch(s6, 64, 7, None, devices=[0]), np.stack([O.kdline(p, 64, 7, 0) for p in s6]))
with capi.tuning(group=0):
    a = synth.uniform(13, 30000, 3)
    check("async cluster kernel", capi.kdline(a, 64, 7, 0), O.kdline(a, 64, 7, 0))
check("vanilla cluster kernel", capi.vanilla(pcs[0], 32, 0), O.fps_vanilla(pcs[0], 32, 0))
check("kd tree", capi.kdtree(pcs[1], 32, 0), O.kdtree(pcs[1], 32, 0))
check("npdu (index window)", capi.npdu(pcs[2], 48, 64, 0), O.fps_npdu(pcs[2], 48, 64, 0))
check("npdu (k nearest)", capi.npdu_kdtree(pcs[3], 48, 32, 0), O.fps_npdu_kdtree(pcs[3], 48, 32, 0))
one = synth.lidar(3, 300000)
check("whole-GPU grid kernel", capi.kdline(one, 128, 9, 0), O.kdline(one, 128, 9, 0))
print("all samplers agree with the oracle")
