import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from fpsample_b200 import capi
B, n, d, k, h, gen, seed, desc = bench.WORKLOADS["cfg2"]
host = np.stack([bench.make_cloud(gen, seed + b, n, d) for b in range(B)])
dp = torch.from_numpy(host).cuda(); do = torch.empty((B, k), dtype=torch.int64, device="cuda")
wsb = capi.workspace_bytes(capi.ALGO_KDLINE, B, n, d, k, h); ws = torch.empty(wsb + 512, dtype=torch.uint8, device="cuda"); wp = (ws.data_ptr() + 255) & ~255
st = torch.cuda.current_stream()
capi.phase_timing(True)
for _ in range(3):
    capi.kdline_batch_dev(dp.data_ptr(), B, n, d, k, 0, h, do.data_ptr(), wp, wsb, st.cuda_stream); print(capi.last_phase_ms())
os.environ["FPS_B200_DBG_BUILD"] = "1"
o = np.zeros(16, dtype=np.uint64); capi.lib().fps_b200_debug_counters(o.ctypes.data)
print("build dbg (cloud 0) cycles: P1 split+chain %d | P2 count %d | P3 rank %d | P4 swap %d | P5 boxes %d | all levels %d" % tuple(int(x) for x in o[1:7]), [int(x) for x in o[:10]])
print(capi.last_plan())
