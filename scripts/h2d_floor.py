"""Host<->device copy floor of this box (no kernels): pinned and pageable H2D, pinned D2H, GB/s.
usage: python scripts/h2d_floor.py [MB]   (under torchrun: every rank copies to its own GPU concurrently)"""
import os, sys, time
import numpy as np, torch
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 512
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
n = mb << 20
pin = torch.empty(n, dtype=torch.uint8).pin_memory()
pag = torch.from_numpy(np.ones(n, dtype=np.uint8))
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return n / best / 1e9
print(f"rank {local}: {mb} MB  pinned H2D {t(lambda: dev.copy_(pin, non_blocking=True)):.1f} GB/s  pageable H2D {t(lambda: dev.copy_(pag)):.1f} GB/s  "
      f"pinned D2H {t(lambda: pin.copy_(dev, non_blocking=True)):.1f} GB/s  host memcpy {t(lambda: pin.copy_(pag)):.1f} GB/s", flush=True)
