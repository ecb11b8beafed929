import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from fpsample_b200 import capi, synth
for B, n, d, k, h in [(2, 200000, 3, 8192, 9), (1, 150000, 3, 8192, 8), (4, 250000, 3, 4096, 9), (1, 200000, 6, 4096, 8)]:
    host = np.stack([synth.uniform(77 + b, n, d) for b in range(B)])
    dp = torch.from_numpy(host).cuda(); do = torch.empty((B, k), dtype=torch.int64, device="cuda")
    res = []
    for mode in (None, "1"):
        if mode: os.environ["FPS_B200_GRID"] = mode
        else: os.environ.pop("FPS_B200_GRID", None)
        wsb = capi.workspace_bytes(capi.ALGO_KDLINE, B, n, d, k, h); ws = torch.empty(wsb + 512, dtype=torch.uint8, device="cuda"); wp = (ws.data_ptr() + 255) & ~255
        st = torch.cuda.current_stream()
        fn = lambda: capi.kdline_batch_dev(dp.data_ptr(), B, n, d, k, 0, h, do.data_ptr(), wp, wsb, st.cuda_stream)
        fn(); torch.cuda.synchronize(); ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        res.append((min(ts), capi.last_plan().split(" + ")[-1][:50], do.cpu().numpy().copy()))
    print(f"B={B} n={n} d={d} k={k} h={h}: default {res[0][0]:.3f} ms [{res[0][1]}] | merged grid {res[1][0]:.3f} ms [{res[1][1]}] same={np.array_equal(res[0][2], res[1][2])}", flush=True)
