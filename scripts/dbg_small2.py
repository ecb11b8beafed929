"""debug: read the region the small-cloud build kernel wrote (single cloud) and compare perm / slots / boxes with the oracle"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from fpsample_b200 import capi, synth
from oracle import oracle
spec = sys.argv[1:] or ["grid", "12", "3000", "2", "6"]
kind, seed, n, d, h = spec[0], int(spec[1]), int(spec[2]), int(spec[3]), int(spec[4])
pc = synth.grid_ties(seed, n, d) if kind == "grid" else synth.uniform(seed, n, d)
k = 10
dp = torch.from_numpy(pc).cuda(); do = torch.empty((1, k), dtype=torch.int64, device="cuda")
wsb = capi.workspace_bytes(capi.ALGO_KDLINE, 1, n, d, k, h); ws = torch.zeros(wsb + 512, dtype=torch.uint8, device="cuda"); off = (-ws.data_ptr()) % 256
capi.kdline_batch_dev(dp.data_ptr(), 1, n, d, k, 0, h, do.data_ptr(), ws.data_ptr() + off, wsb, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize(); print(capi.last_plan())
npad = (n + 31) & ~31; S = 1 << h
pl_ws = 256 + (((npad * 4) + 255) & ~255)
roff = (pl_ws + 255) & ~255
raw = ws[off + roff:].cpu().numpy()
perm = raw[(d + 1) * npad * 4:(d + 2) * npad * 4].view(np.uint32)[:n]
nlo = raw[(d + 2) * npad * 4:(d + 2) * npad * 4 + (S + 1) * 4].view(np.uint32)
nlo_pad = (S + 1 + 31) & ~31
box = raw[((d + 2) * npad + nlo_pad) * 4:((d + 2) * npad + nlo_pad + S * 2 * d) * 4].view(np.float32).reshape(S, 2, d)
operm, olo, obox = oracle.kdline_build(pc, h)
print("perm is a permutation:", np.array_equal(np.sort(perm), np.arange(n)))
bad = np.nonzero(perm != operm)[0]
print("perm mismatches:", bad.size, bad[:20])
# slots: compare non-empty slot ranges in order
mine = [(int(nlo[s]), int(nlo[s + 1])) for s in range(S) if nlo[s + 1] > nlo[s]]
print("oracle leaf_lo:", np.asarray(olo)[:12], "... mine non-empty:", mine[:8], len(mine))
ol = np.asarray(olo)
theirs = [(int(ol[i]), int(ol[i + 1])) for i in range(len(ol) - 1) if ol[i + 1] > ol[i]]
print("ranges equal:", mine == theirs, len(theirs))
if mine != theirs:
    for a, b in zip(mine, theirs):
        if a != b: print("first diff", a, b); break
os.makedirs("gpurun_out", exist_ok=True)
np.savez(f"gpurun_out/dbg_small_{kind}{seed}_{n}_{d}_{h}.npz", perm=perm, nlo=nlo, box=box, operm=operm, olo=np.asarray(olo), obox=obox, pc=pc)
