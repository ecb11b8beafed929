"""cfg5-shaped batches (100k points -> 8192, h=7): the streaming sampler with 1 / 2 / 4 warps per cloud at the per-GPU batch
sizes of a 1..8-GPU split; build / sampling split, executed-work counters, a few clouds checked against the oracle.
usage: python scripts/cmp_cfg5.py D B [B ...] [--wpc=1,2,4] [--check=N] [--split=0|1|2]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fpsample_b200 import capi, synth
from oracle import oracle as O
args = [x for x in sys.argv[1:] if not x.startswith("--")]
opt = dict(x[2:].split("=") for x in sys.argv[1:] if x.startswith("--"))
n, d, k, h = 100000, int(args[0]) if args else 3, 8192, 7
Bs = [int(x) for x in args[1:]] or [512]
wpcs = [int(x) for x in opt.get("wpc", "-1").split(",")]
ncheck = int(opt.get("check", "2"))
if "prefetch" in opt: capi.set_tuning("PREFETCH", int(opt["prefetch"]))
if "split" in opt: capi.set_tuning("STREAM_SPLIT", int(opt["split"]))
nreal = 32
base = np.stack([synth.uniform(3000 + b, n, d) for b in range(nreal)])
want = {b: O.kdline(base[b], k, h, 0) for b in range(ncheck)}
for B in Bs:
    host = np.concatenate([base] * ((B + nreal - 1) // nreal))[:B]
    dp = torch.from_numpy(host).cuda(); do = torch.empty((B, k), dtype=torch.int64, device="cuda")
    for wpc in wpcs:
        capi.set_tuning("STREAM_WARPS", wpc); capi.set_tuning("WARP_GLOBAL_MINB", 1); capi.set_tuning("GROUP", 0)
        wsb = capi.workspace_bytes(capi.ALGO_KDLINE, B, n, d, k, h); ws = torch.empty(wsb + 512, dtype=torch.uint8, device="cuda"); wp = (ws.data_ptr() + 255) & ~255
        st = torch.cuda.current_stream().cuda_stream
        capi.phase_timing(True); best = None
        for _ in range(2):
            capi.kdline_batch_dev(dp.data_ptr(), B, n, d, k, 0, h, do.data_ptr(), wp, wsb, st)
            ph = capi.last_phase_ms()
            if best is None or sum(ph) < sum(best): best = ph
        capi.phase_timing(False)
        capi.set_tuning("COUNT", 1)
        capi.kdline_batch_dev(dp.data_ptr(), B, n, d, k, 0, h, do.data_ptr(), wp, wsb, st)
        cnt = capi.debug_counters(capi.DBG_STREAM)
        capi.set_tuning("COUNT", -1)
        got = do.cpu().numpy().astype(np.uint64)
        ok = all(np.array_equal(got[b + nreal * j], want[b]) for b in want for j in range((B - b + nreal - 1) // nreal) if b + nreal * j < B)
        ok = ok and all(np.array_equal(got[b], got[b % nreal]) for b in range(nreal, B))   # every copy of a cloud, whatever team took it
        pts, pu, fl, early, tests, picks, clouds, stored = [int(x) for x in cnt[:8]]
        byt = pts * 4 * (d + 1) + stored * 4
        print(f"B={B} d={d} wpc={wpc}: build {best[0]:7.2f} ms sampling {best[1]:7.2f} ms -> {B / sum(best) * 1e3:7.0f} clouds/s | parity {'OK' if ok else 'MISMATCH'} | "
              f"per pick: {pts / max(picks, 1):.0f} pts scanned, {pu / max(picks, 1):.0f} point-updates, {fl / max(picks, 1):.2f} passes ({early / max(picks, 1):.2f} early) | "
              f"{byt / 1e9:.1f} GB algorithmic -> {byt / best[1] / 1e6:.0f} GB/s | {capi.last_plan().split('): ')[-1][:150]}", flush=True)
        del ws
