"""time fps_sampling (batched, device-resident) with and without the kd-permutation route (FPS_B200_VANILLA_KD)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fpsample_b200 import capi, synth
shapes = [(1, 8192, 3, 2048), (64, 8192, 3, 2048), (512, 8192, 3, 512), (1, 16384, 3, 4096), (256, 16384, 3, 1024), (32, 50000, 6, 2000),
          (1, 2**20, 3, 8192), (1, 2**20, 3, 65536)]
for B, n, d, k in shapes:
    host = np.stack([synth.uniform(77 + b, n, d) for b in range(B)])
    dp = torch.from_numpy(host).cuda()
    do = torch.empty((B, k), dtype=torch.int64, device="cuda")
    res = []
    for mode in ("0", "1"):
        if mode == "0" and n * k > 3e10:
            res.append((float("nan"), "skipped (minutes)", None)); continue
        capi.set_tuning("VANILLA_KD", int(mode))
        wsb = capi.workspace_bytes(capi.ALGO_VANILLA, B, n, d, k, 0)
        ws = torch.empty(wsb + 512, dtype=torch.uint8, device="cuda")
        wp = (ws.data_ptr() + 255) & ~255
        st = torch.cuda.current_stream()
        fn = lambda: capi.vanilla_batch_dev(dp.data_ptr(), B, n, d, k, 0, do.data_ptr(), wp, wsb, st.cuda_stream)
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        res.append((min(ts), capi.last_plan().split(": ")[-1][-60:], do.cpu().numpy().copy()))
    same = res[0][2] is None or np.array_equal(res[0][2], res[1][2])
    print(f"B={B:4d} n={n:7d} d={d} k={k:5d}: brute force {res[0][0]:9.3f} ms | kd route {res[1][0]:8.3f} ms  same={same}", flush=True)
