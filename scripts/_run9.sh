nvidia-smi -L
timeout 900 python -m pytest tests/test_dist.py tests/test_gpu_parity.py -x -q -m gpu -k "nccl or multi_device or cfg5_production" 2>&1 | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2h_bench2.json 2> gpurun_out/r2h_bench2.err; tail -5 gpurun_out/r2h_bench2.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2h_bench2.json').read().strip().splitlines()[-1])
print("value", l["value"], "e2e", l["e2e"]["value"], l["e2e"]["ms_per_step"], "pageable", l["e2e_pageable"]["ms_per_step"], "floor", l["h2d_floor"]["gb_per_s_per_gpu"], l["h2d_floor"]["e2e_over_floor"])
for w in ("cfg5d3","cfg5d6"):
    e=l["extra"][w]; print(w, "value", e.get("value"), "ms", e.get("ms_per_step"), "e2e", e.get("e2e",{}).get("value"), "e2e ms", e.get("e2e",{}).get("ms_per_step"), e.get("phases_ms"), e.get("error"))
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-300
