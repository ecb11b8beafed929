timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cfg5 or executed_work or gpu_resident or device_pointer or concurrent" 2>&1 | tail -8
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -3 gpurun_out/r2f_bench.err; python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2f_bench.json').read().strip().splitlines()[-1])
def show(d, ind=0):
    for k,v in d.items():
        if isinstance(v, dict): print(' '*ind+k+':'); show(v, ind+2)
        else: print(' '*ind+f"{k}: {str(v)[:230]}")
show(l)
PY
