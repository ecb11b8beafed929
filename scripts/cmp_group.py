"""time the kd-line batch entry with and without the grouped grid sampler (FPS_B200_GROUP) on medium-cloud shapes"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fpsample_b200 import capi, synth
shapes = [(1, 16384, 3, 4096, 7), (8, 50000, 3, 2000, 7), (200, 16384, 3, 1024, 7), (16, 20000, 6, 1000, 6), (4, 98000, 3, 8192, 7),
          (64, 16384, 3, 4096, 7), (32, 8192, 3, 2048, 6), (512, 10000, 3, 512, 6), (2, 200000, 3, 8192, 9)]
for B, n, d, k, h in shapes:
    host = np.stack([synth.uniform(77 + b, n, d) for b in range(B)])
    dp = torch.from_numpy(host).cuda()
    do = torch.empty((B, k), dtype=torch.int64, device="cuda")
    res = []
    for mode in ("0", "1"):
        capi.set_tuning("GROUP", int(mode))
        wsb = capi.workspace_bytes(capi.ALGO_KDLINE, B, n, d, k, h)
        ws = torch.empty(wsb + 512, dtype=torch.uint8, device="cuda")
        wp = (ws.data_ptr() + 255) & ~255
        st = torch.cuda.current_stream()
        fn = lambda: capi.kdline_batch_dev(dp.data_ptr(), B, n, d, k, 0, h, do.data_ptr(), wp, wsb, st.cuda_stream)
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        res.append((min(ts), capi.last_plan().split(" + ")[-1][:46], do.cpu().numpy().copy()))
    same = np.array_equal(res[0][2], res[1][2])
    print(f"B={B:4d} n={n:6d} d={d} k={k:5d} h={h}: default {res[0][0]:8.3f} ms [{res[0][1]}] | group {res[1][0]:8.3f} ms [{res[1][1]}] same={same}", flush=True)
