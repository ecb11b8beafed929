"""time the kd-line build alone (fps_b200_kdline_build_dev): usage python scripts/time_build.py B n d h"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fpsample_b200 import capi, synth
B, n, d, h = [int(x) for x in sys.argv[1:5]]
S = 1 << h
host = synth.uniform_batch(1000, B, n, d)
dp = torch.from_numpy(host).cuda()
perm = torch.empty((B, n), dtype=torch.int32, device="cuda")
lo = torch.empty((B, S + 1), dtype=torch.int32, device="cuda")
box = torch.empty((B, S * 2 * d), dtype=torch.float32, device="cuda")
wsb = capi.workspace_bytes(capi.ALGO_KDLINE, B, n, d, 1, h)
ws = torch.empty(wsb + 512, dtype=torch.uint8, device="cuda")
wp = (ws.data_ptr() + 255) & ~255
st = torch.cuda.current_stream()
fn = lambda: capi.kdline_build_dev(dp.data_ptr(), B, n, d, h, perm.data_ptr(), lo.data_ptr(), box.data_ptr(), wp, wsb, st.cuda_stream)
fn(); torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(f"build only B={B} n={n} d={d} h={h}: min {min(ts):.3f} ms | {capi.last_plan()}")
os.environ["FPS_B200_DBG_BUILD"] = "1"
o = np.zeros(16, dtype=np.uint64); capi.lib().fps_b200_debug_counters(o.ctypes.data)
print("  build dbg (cloud 0) cycles: P1 split+chain %d | P2 count %d | P3 rank %d | P4 swap %d | P5 boxes %d | all levels %d" % tuple(int(x) for x in o[1:7]))
