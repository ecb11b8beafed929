"""cfg5-shaped batches (100k points -> 8192, h=7) at the per-GPU batch sizes of an 8-GPU shard: grouped grid sampler
(planner default below 4 clouds per SM) against the one-warp-per-cloud streaming kernel; build / sampling split."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from fpsample_b200 import capi, synth
n, d, k, h = 100000, int(sys.argv[1]) if len(sys.argv) > 1 else 3, 8192, 7
Bs = [int(x) for x in sys.argv[2:]] or [148, 296, 512]
base = np.stack([synth.uniform(3000 + b, n, d) for b in range(16)])
for B in Bs:
    host = np.concatenate([base] * ((B + 15) // 16))[:B]
    host = host + (np.arange(B, dtype=np.float32) * 1e-3)[:, None, None]
    dp = torch.from_numpy(host).cuda(); do = torch.empty((B, k), dtype=torch.int64, device="cuda")
    res = {}
    for name, env in (("default", {}), ("warpg", {"FPS_B200_WARP_GLOBAL_MINB": "1", "FPS_B200_GROUP": "0"})):
        for kk in ("FPS_B200_WARP_GLOBAL_MINB", "FPS_B200_GROUP"): os.environ.pop(kk, None)
        os.environ.update(env)
        wsb = capi.workspace_bytes(capi.ALGO_KDLINE, B, n, d, k, h); ws = torch.empty(wsb + 512, dtype=torch.uint8, device="cuda"); wp = (ws.data_ptr() + 255) & ~255
        capi.phase_timing(True)
        best = None
        for _ in range(3):
            capi.kdline_batch_dev(dp.data_ptr(), B, n, d, k, 0, h, do.data_ptr(), wp, wsb, torch.cuda.current_stream().cuda_stream)
            ph = capi.last_phase_ms()
            if best is None or sum(ph) < sum(best): best = ph
        capi.phase_timing(False)
        res[name] = (best, do.cpu().numpy().copy(), capi.last_plan())
        del ws
    same = np.array_equal(res["default"][1], res["warpg"][1])
    for name in res:
        b, _, plan = res[name]
        print(f"B={B} d={d} {name:8s}: build {b[0]:8.2f} ms sampling {b[1]:8.2f} ms -> {B / (b[0] + b[1]) * 1e3:8.0f} clouds/s | {plan[:150]}")
    print("   same indices:", same)
