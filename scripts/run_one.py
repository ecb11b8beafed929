"""Run one workload a few times through the device-pointer C ABI (for ncu launch lists / quick timings).
usage: python scripts/run_one.py <workload> <algo> [reps] [B override]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from fpsample_b200 import capi

wl, algo = sys.argv[1], sys.argv[2]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
B, n, d, k, h, gen, seed, desc = bench.WORKLOADS[wl]
if len(sys.argv) > 4: B = int(sys.argv[4])
host = np.stack([bench.make_cloud(gen, seed + b, n, d) for b in range(B)])
dp = torch.from_numpy(host).cuda()
do = torch.empty((B, k), dtype=torch.int64, device="cuda")
a = capi.ALGO_VANILLA if algo == "vanilla" else capi.ALGO_KDLINE
wsb = capi.workspace_bytes(a, B, n, d, k, h)
ws = torch.empty(wsb + 512, dtype=torch.uint8, device="cuda")
wp = (ws.data_ptr() + 255) & ~255
st = torch.cuda.current_stream()
def fn():
    if algo == "vanilla": capi.vanilla_batch_dev(dp.data_ptr(), B, n, d, k, 0, do.data_ptr(), wp, wsb, st.cuda_stream)
    else: capi.kdline_batch_dev(dp.data_ptr(), B, n, d, k, 0, h, do.data_ptr(), wp, wsb, st.cuda_stream)
fn(); torch.cuda.synchronize()
ts = []
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(f"{wl} {algo} B={B}: min {min(ts):.3f} ms mean {sum(ts)/len(ts):.3f} ms -> {B/min(ts)*1e3:.1f} clouds/s | {capi.last_plan()}")
if algo == "kdline" and "dist" in capi.last_plan():
    o = capi.debug_counters(capi.DBG_ASYNC)
    it = max(int(o[0]), 1)
    if int(o[0]): print("  dist dbg (cluster 0, CTA 0): iterations %d picks/iter %.2f | per iteration: cand+send %.0f wait %.0f select %.0f tests %.0f flush %.0f = %.0f cyc | flushed buckets/iter (this CTA) %.2f items/iter %.2f" % (
        o[0], o[1] / it, o[2] / it, o[3] / it, o[4] / it, o[5] / it, o[6] / it, sum(int(x) for x in o[2:7]) / it, o[7] / it, o[8] / it))
if algo == "kdline" and "kdline_grid" in capi.last_plan():
    o = capi.debug_counters(capi.DBG_GRID)
    it = max(int(o[0]), 1)
    print("  grid dbg (CTA 0): rounds %d picks/round %.1f | cycles per round: apply+select %.0f merge+publish %.0f gather %.0f eligible+rank %.0f pairs %.0f out+rel %.0f = %.0f" % (
        o[0], o[1] / it, o[2] / it, o[3] / it, o[4] / it, o[5] / it, o[6] / it, o[7] / it, sum(int(x) for x in o[2:8]) / it))
    print("     eligible+rank split: bounds %.0f compaction %.0f rank %.0f table %.0f | E/round %.1f | picks phase: conflict matrix %.0f fixed point %.0f (%.1f iterations) floors %.0f" % (o[8] / it, o[9] / it, o[10] / it, 0.0, o[12] / it, o[13] / it, o[14] / it, o[11] / it, o[15] / it))
if algo == "kdline" and "async" in capi.last_plan():
    d = capi.debug_counters(); it = max(d["iterations"], 1)
    print("  dbg:", d, "| per iteration:", {k: round(v / it, 1) for k, v in d.items() if k.startswith("cyc")}, "picks/iter %.2f" % (d["picks"] / it))
if algo == "kdline" and "warp" in capi.last_plan():
    out = capi.debug_counters(capi.DBG_WARP)
    it = max(int(out[0]), 1)
    if int(out[0]): print("  warp dbg (cloud 0): picks %d | per pick: test %.0f scan %.0f reduce %.0f argmax %.0f total %.0f cyc | buckets/pick %.2f groups/pick %.2f" % (
        out[0], out[1] / it, out[2] / it, out[3] / it, out[4] / it, out[7] / it, out[5] / it, out[6] / it))
