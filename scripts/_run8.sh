for v in "" lb256 lb256p lb256n6 p512; do
  echo "=== variant: ${v:-default}"
  if [ -n "$v" ]; then export FPS_B200_LIB=$PWD/fpsample_b200/variants/libfps_$v.so; fi
  python scripts/run_one.py cfg2 kdline 5 2>&1 | tail -2 | cut -c1-200
  python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch, bench
from fpsample_b200 import capi
B, n, d, k, h, gen, seed, desc = bench.WORKLOADS["cfg2"]
host = np.stack([bench.make_cloud(gen, seed + b, n, d) for b in range(B)])
dp = torch.from_numpy(host).cuda(); do = torch.empty((B, k), dtype=torch.int64, device="cuda")
wsb = capi.workspace_bytes(capi.ALGO_KDLINE, B, n, d, k, h); ws = torch.empty(wsb + 512, dtype=torch.uint8, device="cuda"); wp = (ws.data_ptr() + 255) & ~255
st = torch.cuda.current_stream().cuda_stream
capi.phase_timing(True); best = None
for _ in range(6):
    capi.kdline_batch_dev(dp.data_ptr(), B, n, d, k, 0, h, do.data_ptr(), wp, wsb, st); ph = capi.last_phase_ms()
    if best is None or ph[1] < best[1]: best = ph
from oracle import oracle as O
ok = all(np.array_equal(do[b].cpu().numpy().astype(np.uint64), O.kdline(host[b], k, h, 0)) for b in (0, 500, 1023))
print(f"   build {best[0]:.3f} ms sampling {best[1]:.3f} ms parity {ok}")
PY
done
# pageable e2e check + cfg5 e2e
unset FPS_B200_LIB
python bench.py --steps 5 --warmup 3 --no-extras --no-cpu > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; tail -2 gpurun_out/r2g_bench.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2g_bench.json').read().strip().splitlines()[-1])
print("e2e", l["e2e"]["ms_per_step"], "pageable", l["e2e_pageable"]["ms_per_step"], "floor", l["h2d_floor"])
for w in ("cfg5d3","cfg5d6"):
    e=l["extra"][w]; print(w, "value", e.get("value"), "ms", e.get("ms_per_step"), "e2e ms", e.get("e2e",{}).get("ms_per_step"), e.get("error"))
PY
