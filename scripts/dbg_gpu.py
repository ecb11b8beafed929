"""Ad-hoc GPU bring-up script (not a pytest file): compares the CUDA path with the oracle, verbosely."""
import os, sys, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fpsample_b200 import capi, synth
from oracle import oracle as O

def cmp(name, got, want):
    ok = np.array_equal(got, want)
    if ok:
        print(f"PASS {name}", flush=True)
    else:
        bad = np.argwhere(got != want)
        print(f"FAIL {name}: {len(bad)} mismatches, first at {bad[0]}, got {got.ravel()[:8]} want {want.ravel()[:8]}", flush=True)
    return ok

def run(name, fn):
    t = time.time()
    try:
        r = fn()
        print(f"  [{name}] {time.time()-t:.3f}s plan={capi.last_plan()}", flush=True)
        return r
    except Exception as e:
        print(f"ERROR {name}: {type(e).__name__}: {e}", flush=True)
        return None

print("devices", capi.device_count(), capi.lib().fps_b200_version().decode(), flush=True)
allok = True
# vanilla single cloud, several shapes/dims
for (n, d, k, s) in [(10, 3, 5, 2), (1000, 3, 100, 7), (4096, 3, 1024, 0), (4096, 6, 512, 5), (5000, 2, 300, 1), (3000, 8, 200, 0), (16384, 3, 1024, 3), (777, 1, 200, 4), (2048, 5, 100, 9), (1500, 7, 100, 9), (40000, 3, 300, 11), (3000, 12, 100, 2)]:
    pc = synth.uniform(n + d, n, d)
    got = run(f"vanilla n={n} d={d} k={k}", lambda: capi.vanilla(pc, k, s))
    if got is not None: allok &= cmp(f"vanilla n={n} d={d} k={k}", got, O.fps_vanilla(pc, k, s))
    else: allok = False
# ties
for seed in range(3):
    g = synth.grid_ties(seed, 3000, 3)
    got = run("vanilla ties", lambda: capi.vanilla(g, 500, [5, 1, 9]))
    if got is not None: allok &= cmp(f"vanilla ties multi-start seed={seed}", got, O.fps_vanilla(g, 500, [5, 1, 9]))
    else: allok = False
# big vanilla: cluster >1 and grid
for (n, d, k) in [(100000, 3, 500), (100000, 6, 300), (300000, 3, 200)]:
    pc = synth.uniform(n, n, d)
    got = run(f"vanilla big n={n} d={d}", lambda: capi.vanilla(pc, k, 0))
    if got is not None: allok &= cmp(f"vanilla big n={n} d={d}", got, O.fps_vanilla(pc, k, 0))
    else: allok = False
# vanilla batch
pcs = synth.uniform_batch(1000, 37, 4096, 3)
got = run("vanilla batch", lambda: capi.vanilla_batch(pcs, 256, np.arange(37)))
if got is not None: allok &= cmp("vanilla batch 37x4096", got, np.stack([O.fps_vanilla(pcs[b], 256, b) for b in range(37)]))
else: allok = False
# kdline
for (n, d, k, h, s) in [(64, 3, 20, 2, 1), (1000, 3, 100, 3, 7), (4096, 3, 1024, 5, 0), (4096, 3, 1024, 7, 3), (4096, 6, 512, 5, 5), (5000, 2, 300, 4, 1), (3000, 8, 200, 6, 0), (16384, 3, 4096, 7, 0), (777, 1, 200, 3, 4), (4096, 3, 200, 12, 0), (100000, 3, 2000, 7, 0), (100000, 6, 1000, 9, 0)]:
    pc = synth.uniform(n + d + h, n, d)
    got = run(f"kdline n={n} d={d} k={k} h={h}", lambda: capi.kdline(pc, k, h, s))
    if got is not None: allok &= cmp(f"kdline n={n} d={d} k={k} h={h}", got, O.kdline(pc, k, h, s))
    else: allok = False
for seed in range(3):
    g = synth.grid_ties(seed, 3000, 3)
    got = run("kdline ties", lambda: capi.kdline(g, 500, 6, seed))
    if got is not None: allok &= cmp(f"kdline ties seed={seed}", got, O.kdline(g, 500, 6, seed))
    else: allok = False
pcs = synth.uniform_batch(1000, 300, 4096, 3)
got = run("kdline batch", lambda: capi.kdline_batch(pcs, 1024, 5))
if got is not None: allok &= cmp("kdline batch 300x4096", got, np.stack([O.kdline(pcs[b], 1024, 5, 0) for b in range(300)]))
else: allok = False
if "--big" in sys.argv:
    pc = synth.uniform(5, 2**20, 3)
    got = run("kdline 1M", lambda: capi.kdline(pc, 65536, 9, 0))
    if got is not None: allok &= cmp("kdline 1M", got, O.kdline(pc, 65536, 9, 0))
print("ALL OK" if allok else "SOME FAILED", "launches", capi.kernel_launches(), flush=True)
