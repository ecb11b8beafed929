"""debug: kdline through the small-cloud build kernel vs the general build kernel (FPS_B200_KDSMALL=0), case by case"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from fpsample_b200 import capi, synth
cases = [("grid", 11, 3000, 1, 6), ("grid", 12, 3000, 2, 6), ("grid", 13, 3000, 3, 6), ("grid", 16, 3000, 6, 6), ("grid", 99, 64, 2, 3), ("uni", 42, 2000, 2, 4), ("uni", 45, 2000, 5, 4), ("uni", 1, 4099, 3, 5)]
only = sys.argv[1:] 
for c in cases:
    if c[0] == "grid": pc = synth.grid_ties(c[1], c[2], c[3], *( [3] if c[1]==99 else [])); 
    else: pc = synth.uniform(c[1], c[2], c[3])
    h = c[4] if c[0] != "grid" else 6
    k = min(500, pc.shape[0])
    os.environ["FPS_B200_KDSMALL"] = "1"; a = capi.kdline(pc, k, h, 0); pa = capi.last_plan()
    os.environ["FPS_B200_KDSMALL"] = "0"; b = capi.kdline(pc, k, h, 0)
    bad = np.nonzero(a != b)[0]
    print(c, "same" if bad.size == 0 else f"DIFF first at {bad[0]} of {k}: {a[bad[0]]} vs {b[bad[0]]} ({bad.size} differ)", "|", pa[:60])
