"""cfg5-shaped batch: per-cloud CTA build (kdline_kernel, data in L2) against the grid-wide gb_* launches"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from fpsample_b200 import capi, synth
n, d, k, h = 100000, int(sys.argv[1]) if len(sys.argv) > 1 else 3, 8192, 7
for B in [int(x) for x in sys.argv[2:]] or [512]:
    base = np.stack([synth.uniform(3000 + b, n, d) for b in range(8)])
    host = np.concatenate([base] * ((B + 7) // 8))[:B] + (np.arange(B, dtype=np.float32) * 1e-3)[:, None, None]
    dp = torch.from_numpy(host).cuda(); do = torch.empty((B, k), dtype=torch.int64, device="cuda")
    res = {}
    for name, env in (("cta build", "0"), ("grid build", "1")):
        capi.set_tuning("GRIDBUILD", int(env))
        wsb = capi.workspace_bytes(capi.ALGO_KDLINE, B, n, d, k, h); ws = torch.empty(wsb + 512, dtype=torch.uint8, device="cuda"); wp = (ws.data_ptr() + 255) & ~255
        capi.phase_timing(True); best = None
        for _ in range(3):
            capi.kdline_batch_dev(dp.data_ptr(), B, n, d, k, 0, h, do.data_ptr(), wp, wsb, torch.cuda.current_stream().cuda_stream)
            ph = capi.last_phase_ms()
            if best is None or sum(ph) < sum(best): best = ph
        capi.phase_timing(False)
        res[name] = do.cpu().numpy().copy()
        print(f"B={B} d={d} {name:10s}: build {best[0]:8.2f} ms sampling {best[1]:8.2f} ms ws {wsb >> 20} MiB | {capi.last_plan()[:90]}")
        del ws
    print("   same:", np.array_equal(res["cta build"], res["grid build"]))
