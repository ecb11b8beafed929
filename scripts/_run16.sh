timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python - <<'PY'
import time, numpy as np, sys, os
sys.path.insert(0, os.getcwd())
import fpsample_b200 as fps
from fpsample_b200 import capi, synth
from oracle import oracle as O
ref = O.load_reference()
for n, k in ((4096, 1024), (100000, 8192)):
    pc = synth.uniform(1, n, 3)
    fps.fps_npdu_kdtree_sampling(pc, k, start_idx=0)
    t = time.perf_counter(); a = fps.fps_npdu_kdtree_sampling(pc, k, start_idx=0); tg = time.perf_counter() - t
    t = time.perf_counter(); b = ref.fps_npdu_kdtree_sampling(pc, k, start_idx=0); tr = time.perf_counter() - t
    print(f"npdu_kdtree {n}->{k}: gpu {tg*1e3:.2f} ms, reference on one core {tr*1e3:.2f} ms, equal {np.array_equal(a, b)}")
PY
