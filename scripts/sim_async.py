"""CPU simulation of the asynchronous coordinator/worker pick protocol (design validation, not product code).
Buckets are EXACT (max known, pending refs do not touch the max point) or INFLIGHT (a scan job is out, only an
upper bound is known).  Jobs complete after a random number of picks.  Output must equal the oracle's."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
from fpsample_b200 import synth

f32 = np.float32
FLT_MAX = np.finfo(np.float32).max

def sqd(a, b):
    acc = f32(0)
    for j in range(len(a)):
        t = f32(a[j] - b[j]); acc = f32(acc + f32(t * t))
    return acc

def boxd(r, lo, hi):
    acc = f32(0)
    for j in range(len(r)):
        e = f32(0)
        if r[j] > hi[j]: e = f32(r[j] - hi[j])
        elif r[j] < lo[j]: e = f32(lo[j] - r[j])
        acc = f32(acc + f32(e * e))
    return acc

def simulate(pc, k, h, start, rng, max_delay=6, R=4):
    perm, bounds, box = O.kdline_build(pc, h)
    q = pc[perm.astype(np.int64)]
    n = len(q); nb = len(bounds) - 1
    dis = np.full(n, FLT_MAX, dtype=np.float32)
    def scan(b, refs):
        lo, hi = int(bounds[b]), int(bounds[b + 1])
        seg = q[lo:hi]
        v = dis[lo:hi]
        for r in refs:
            d = np.zeros(hi - lo, dtype=np.float32)
            for j in range(q.shape[1]):
                t = (seg[:, j] - q[r, j]).astype(np.float32); d = (d + (t * t).astype(np.float32)).astype(np.float32)
            v = np.minimum(v, d)
        dis[lo:hi] = v
        i = int(np.argmax(v))  # first max = lowest position
        mx = v[i]
        snd = np.max(np.delete(v, i)) if hi - lo > 1 else f32(-1)
        return mx, lo + i, snd
    # state
    EX, INF = 0, 1
    state = [EX] * nb; mx = [f32(0)] * nb; pos = [0] * nb; snd = [f32(0)] * nb; U = [f32(0)] * nb
    pend = [[] for _ in range(nb)]; jobs = {}  # b -> (due_time, result)
    stalls = 0; njobs = 0
    # init pass (KDNode::init): every bucket scans the first ref
    for b in range(nb):
        mx[b], pos[b], snd[b] = scan(b, [start])
    out = [int(perm[start])]
    cur = start
    t = 0
    def issue(b, ub):
        nonlocal njobs
        refs = pend[b]; pend[b] = []
        res = scan(b, refs)           # result computed now (dis in memory updated in job order), delivered later
        jobs[b] = (t + rng.randint(1, max_delay), res)
        state[b] = INF; U[b] = ub; njobs += 1
    def integrate(b):
        m, p, s = jobs.pop(b)[1]
        mx[b], pos[b], snd[b] = m, p, s
        keep = []; dirty = False; dmin = None
        for r in pend[b]:
            if boxd(q[r], box[b, 0], box[b, 1]) >= m: continue
            d = sqd(q[p], q[r])
            if d > m: keep.append(r)
            else:
                keep.append(r); dirty = True
                dmin = d if dmin is None else min(dmin, d)
        pend[b] = keep
        if dirty: issue(b, max(s, min(m, dmin)))
        else: state[b] = EX
    for it in range(1, k):
        while True:
            # deliver due results
            for b in [b for b in list(jobs) if jobs[b][0] <= t]: integrate(b)
            best = None
            for b in range(nb):
                key = (mx[b], -pos[b], 0) if state[b] == EX else (U[b], 1, 1)
                if best is None or key > best[0]: best = (key, b)
            if best[0][2] == 0: break
            stalls += 1; t += 1
        c = best[1]; p = pos[c]; out.append(int(perm[p])); t += 1
        for b in range(nb):
            bd = boxd(q[p], box[b, 0], box[b, 1])
            if state[b] == INF:
                if bd < U[b]:
                    pend[b].append(p)
                    assert len(pend[b]) < 64
                continue
            if bd >= mx[b]: continue
            d = sqd(q[pos[b]], q[p])
            pend[b].append(p)
            if d > mx[b]:
                if len(pend[b]) >= R: issue(b, mx[b])      # list full: forced flush
                continue
            issue(b, max(snd[b], min(mx[b], d)))
    return np.array(out, dtype=np.uint64), stalls, njobs

if __name__ == "__main__":
    rng = random.Random(1)
    tot = 0
    for (gen, n, d, k, h, s) in [("u", 2000, 3, 400, 5, 3), ("g", 1500, 2, 500, 5, 0), ("g", 1200, 3, 400, 6, 7), ("u", 3000, 6, 300, 4, 1),
                                 ("l", 4000, 3, 600, 6, 2), ("g", 800, 1, 300, 4, 5), ("u", 4096, 3, 1024, 5, 0), ("d", 50, 3, 20, 3, 2)]:
        pc = {"u": lambda: synth.uniform(n, n, d), "g": lambda: synth.grid_ties(n, n, d, 5), "l": lambda: synth.lidar(n, n),
              "d": lambda: np.full((n, d), 0.5, np.float32)}[gen]()
        for md in (1, 3, 9):
            got, stalls, njobs = simulate(pc, k, h, s, rng, max_delay=md)
            want = O.kdline(pc, k, h, s)
            ok = np.array_equal(got, want)
            print(gen, n, d, k, h, "delay", md, "OK" if ok else "MISMATCH at %d" % np.argmax(got != want), "stalls", stalls, "jobs/pick %.2f" % (njobs / k), flush=True)
            tot += not ok
    print("FAILURES", tot)

def big():
    rng = random.Random(2)
    n, d, k, h = 200000, 3, 1500, 9
    pc = synth.uniform(7, n, d)
    for md in (4, 8, 16):
        got, stalls, njobs = simulate(pc, k, h, 0, rng, max_delay=md, R=12)
        want = O.kdline(pc, k, h, 0)
        print("big delay", md, np.array_equal(got, want), "stalls", stalls, "jobs/pick %.2f" % (njobs / k), flush=True)

def ratio():
    rng = random.Random(3)
    n, d, k, h = 65536, 3, 4096, 6
    pc = synth.uniform(8, n, d)
    for md in (4, 8):
        for kk in (512, 4096):
            got, stalls, njobs = simulate(pc, kk, h, 0, rng, max_delay=md, R=12)
            want = O.kdline(pc, kk, h, 0)
            print("ratio delay", md, "k", kk, np.array_equal(got, want), "stalls", stalls, "jobs/pick %.2f" % (njobs / kk), flush=True)
