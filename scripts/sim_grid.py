"""CPU model of the grid-wide batched pick protocol of kdline_grid.cu (design validation, not product code).

The permuted cloud is cut into contiguous position slices (one warp each, 32 warps per CTA, one CTA per SM).  Every
round each CTA publishes its M largest (value, position) keys plus a BOUND key (everything it did not publish sorts
at or below it); all CTAs then select, redundantly and identically, the longest prefix of the globally sorted
published keys that (i) lies above every CTA's bound and (ii) is not lowered by an earlier pick of the same round
(dist(P_j, P_i) >= val_j for i < j).  That prefix is exactly the next J picks of the sequential recurrence
(SURVEY.md A.4).  Picks are then applied eagerly, pruned by per-slice boxes.  Output must equal the oracle's.

With past_conflicts=True (what the kernel does) a lowered candidate is skipped and only raises the floor later picks
of the round must beat: 1457 -> 531 rounds at 2^20 -> 65536 (uniform), 1334 -> 675 (lidar), still bit-exact.
(The model cuts slices by position; the kernel additionally aligns them to kd leaves, which only tightens the boxes.)

  python scripts/sim_grid.py [n] [k] [h] [gen] [seed] [M] [past_conflicts 0|1]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
from fpsample_b200 import synth

f32 = np.float32
FLT_MAX = np.finfo(np.float32).max


def sqd_rows(seg, r):
    d = None
    for j in range(seg.shape[-1]):
        t = (seg[..., j] - r[j]).astype(f32)
        t2 = (t * t).astype(f32)
        d = t2 if d is None else (d + t2).astype(f32)
    return d


def boxd_rows(lo, hi, r):
    acc = np.zeros(lo.shape[0], dtype=f32)
    for j in range(lo.shape[1]):
        e = np.maximum(np.maximum((r[j] - hi[:, j]).astype(f32), (lo[:, j] - r[j]).astype(f32)), f32(0))
        acc = (acc + (e * e).astype(f32)).astype(f32)
    return acc


def simulate(pc, k, h, start, G=148, W=32, PPT=None, M=8, KW=2, ECAP=256, verbose=True, past_conflicts=False):
    perm, bounds, box = O.kdline_build(pc, h)
    perm = perm.astype(np.int64)
    q = pc[perm]
    n, D = q.shape
    if PPT is None:
        PPT = -(-n // (G * W * 32))
    SL = 32 * PPT
    NS = G * W
    npad = NS * SL
    assert npad >= n
    qp = np.zeros((npad, D), dtype=f32)
    qp[:n] = q
    dis = np.full(npad, -1.0, dtype=f32)   # padding never wins
    dis[:n] = FLT_MAX
    qs = qp.reshape(NS, SL, D)
    vs = dis.reshape(NS, SL)
    valid = (np.arange(npad) < n).reshape(NS, SL)
    lo = np.where(valid[..., None], qs, np.inf).min(axis=1).astype(f32)
    hi = np.where(valid[..., None], qs, -np.inf).max(axis=1).astype(f32)
    empty = ~valid.any(axis=1)
    smax = np.where(empty, f32(-1), FLT_MAX).astype(f32)
    # slice top lists: KW candidates + bound key (u64 keys: value bits << 32 | ~pos)
    def keys_of(s_idx):
        v = vs[s_idx]                                       # [m, SL]
        pos = (s_idx[:, None] * SL + np.arange(SL)[None, :]).astype(np.uint64)
        kb = v.view(np.uint32).astype(np.uint64) << np.uint64(32)
        key = np.where(v >= 0, kb | (np.uint64(0xfffffffe) - pos), np.uint64(0))
        return key
    stop = np.zeros((NS, KW + 1), dtype=np.uint64)
    def reselect(s_idx):
        key = keys_of(s_idx)
        part = -np.sort(-key.view(np.int64), axis=1)[:, :KW + 1]   # keys < 2^63 (value bits of finite floats)
        stop[s_idx] = part.view(np.uint64) if part.dtype == np.uint64 else part.astype(np.uint64)
        smax[s_idx] = vs[s_idx].max(axis=1)
    out = np.empty(k, dtype=np.int64)
    out[0] = start
    t = 1
    acc_pos = [start]
    rounds = 0
    Js, Es, touch_max, dirty_cnt, Ms = [], [], [], [], []
    t0 = time.time()
    while True:
        # ---- apply the accepted picks (eager, pruned by slice boxes) ------------------------------------------
        dirty = np.zeros(NS, dtype=bool)
        tcount = np.zeros(NS, dtype=np.int32)
        for p in acc_pos:
            r = qp[p]
            bd = boxd_rows(lo, hi, r)
            tch = np.nonzero(bd < smax)[0]
            if tch.size:
                d = sqd_rows(qs[tch], r)
                nv = np.minimum(vs[tch], d)
                nv = np.where(valid[tch], nv, f32(-1))
                vs[tch] = nv
                dirty[tch] = True
                tcount[tch] += 1
        di = np.nonzero(dirty)[0]
        if di.size:
            reselect(di)
        touch_max.append(int(tcount.max()))
        dirty_cnt.append(int(dirty.reshape(G, W).sum(axis=1).max()))
        if t >= k:
            break
        # ---- CTA merge: top-M of the warps' candidates, bound = max(M+1-th, warp bounds) ------------------------
        ck = stop[:, :KW].reshape(G, W * KW)
        wb = stop[:, KW].reshape(G, W).max(axis=1)
        srt = -np.sort(-ck.view(np.int64), axis=1)
        srt = srt.astype(np.uint64)
        # ---- selection (identical on every CTA): adaptive M' so that at most ECAP candidates are eligible ----------
        Mu = M
        while True:
            pub = srt[:, :Mu]
            cb = np.maximum(srt[:, Mu], wb)
            Bd = cb.max()
            el = pub[pub > Bd]
            if el.size <= ECAP or Mu == 1:
                break
            Mu //= 2
        el = -np.sort(-el.view(np.int64)).astype(np.int64)
        el = el.astype(np.uint64)
        E = el.size
        assert E >= 1, "no eligible candidate: the protocol would stall"
        vals = (el >> np.uint64(32)).astype(np.uint32).view(f32)
        poss = (np.uint64(0xfffffffe) - (el & np.uint64(0xffffffff))).astype(np.int64)
        P = qp[poss]
        # first candidate lowered by an earlier one
        J = E
        if vals[0] == 0:
            out[t:] = poss[0]
            t = k
            acc_pos = []
            rounds += 1
            continue
        nz = np.nonzero(vals == 0)[0]
        if nz.size:
            J = min(J, int(nz[0]))
        if past_conflicts:
            # continue past a lowered candidate: it only raises the floor every later pick of the round must beat
            floor = Bd
            acc_idx = [0]
            for j in range(1, J):
                if el[j] <= floor:
                    break
                d = sqd_rows(P[acc_idx], P[j])
                m = d.min()
                if m < vals[j]:
                    nk = (np.uint64(np.float32(m).view(np.uint32)) << np.uint64(32)) | (el[j] & np.uint64(0xffffffff))
                    floor = max(floor, nk)
                else:
                    acc_idx.append(j)
                if len(acc_idx) >= k - t:
                    break
            acc_idx = acc_idx[:k - t]
            J = len(acc_idx)
            out[t:t + J] = poss[acc_idx]
            acc_pos = list(poss[acc_idx])
        else:
            for j in range(1, J):
                d = sqd_rows(P[:j], P[j])
                if (d < vals[j]).any():
                    J = j
                    break
            J = min(J, k - t)
            out[t:t + J] = poss[:J]
            acc_pos = list(poss[:J])
        t += J
        rounds += 1
        Js.append(J); Es.append(E); Ms.append(Mu)
        if verbose and rounds % 100 == 0:
            print(f"round {rounds} t={t} J={J} E={E} M'={Mu} maxtouch={touch_max[-1]} dirty/CTA max={dirty_cnt[-1]} ({time.time()-t0:.1f}s)", flush=True)
    ids = perm[out]
    return ids, dict(rounds=rounds, J=np.array(Js), E=np.array(Es), Mu=np.array(Ms), touch_max=np.array(touch_max), dirty=np.array(dirty_cnt), PPT=PPT)


if __name__ == "__main__":
    a = sys.argv[1:]
    n = int(a[0]) if len(a) > 0 else 1 << 20
    k = int(a[1]) if len(a) > 1 else 65536
    h = int(a[2]) if len(a) > 2 else 9
    gen = a[3] if len(a) > 3 else "uniform"
    seed = int(a[4]) if len(a) > 4 else 5
    M = int(a[5]) if len(a) > 5 else 8
    pc = synth.lidar(seed, n) if gen == "lidar" else synth.uniform(seed, n, 3)
    ids, st = simulate(pc, k, h, 0, M=M, past_conflicts=len(a) > 6 and a[6] == '1')
    ref = O.kdline(pc, k, h, 0).astype(np.int64)
    print("equal to oracle:", np.array_equal(ids, ref))
    J, E = st["J"], st["E"]
    print(f"rounds={st['rounds']} PPT={st['PPT']} J mean={J.mean():.1f} median={np.median(J)} max={J.max()}  E mean={E.mean():.1f} max={E.max()}")
    print("J==E (ran out of eligible) rounds:", int((J == E).sum()), " M' histogram:", np.unique(st["Mu"], return_counts=True))
    print("max touches per slice per round: mean", st["touch_max"].mean(), "max", st["touch_max"].max(), " dirty warps per CTA (max over CTAs): mean", st["dirty"].mean())
    q = np.cumsum(J)
    for thr in (1024, 4096, 16384, 65536):
        print(f"  rounds to reach t={thr}: {int(np.searchsorted(q, thr - 1)) + 1}")
