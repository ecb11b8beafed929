import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from fpsample_b200 import capi, synth
for (B, n, k, w) in [(1, 4096, 1024, 64), (1, 100000, 8192, 195), (1024, 4096, 1024, 64)]:
    pcs = synth.uniform_batch(1, B, n, 3)
    capi.npdu_batch(pcs, k, w, None, devices=[0])
    t = time.perf_counter(); capi.npdu_batch(pcs, k, w, None, devices=[0]); dt = time.perf_counter() - t
    print(f"npdu B={B} n={n} k={k} w={w}: {dt*1e3:.2f} ms end to end ({B/dt:.0f} clouds/s)")
