"""small kd-line / kd-tree / vanilla calls for compute-sanitizer (memcheck, racecheck, synccheck)"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from fpsample_b200 import capi, synth
pcs = synth.uniform_batch(5, 6, 2048, 3)
print(capi.kdline_batch(pcs, 64, 5, None, devices=[0])[:, :4].tolist(), capi.last_plan()[:60])
lat = np.stack([synth.grid_ties(40 + b, 1500, 2) for b in range(4)])
print(capi.kdline_batch(lat, 48, 6, None, devices=[0])[:, :4].tolist())
big = synth.uniform_batch(9, 2, 16384, 3)
print(capi.kdline_batch(big, 64, 7, None, devices=[0])[:, :4].tolist(), capi.last_plan()[:60])
print(capi.vanilla(pcs[0], 32, 0)[:4].tolist(), capi.kdtree(pcs[1], 32, 0)[:4].tolist())
one = synth.lidar(3, 300000)
print(capi.kdline(one, 256, 9, 0)[:4].tolist(), capi.last_plan()[:60])
