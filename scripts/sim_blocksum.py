"""CPU model of the exact block-parallel sequential binary32 sum (the kd build's split value, KDTreeBase.h:151-158).

s_{i+1} = RN(s_i + x_i).  While every partial sum stays inside the binade of the incoming sum (|s| in [2^e, 2^(e+1)),
ulp u = 2^(e-23)), each addition is s + RN_u(x) -- an INTEGER addition in units of u -- unless x/u lies exactly half way
between two integers (a tie, decided by the parity of the running sum).  So a block of elements can be summed with an
integer prefix scan; the block is accepted if (a) no element is a tie, (b) every prefix k satisfies 2^23 < |k| < 2^24 with
the sign of the incoming sum.  Otherwise the block is summed sequentially.  This script checks the result bit for bit
against the sequential sum and reports how many blocks fall back."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fpsample_b200 import synth

def seq_sum(x):
    s = np.float32(0)
    for v in x: s = np.float32(s + v)
    return s

def block_sum(x, blk=1024):
    s = np.float32(0); i = 0; n = len(x); fast = slow = 0
    while i < n:
        m = min(blk, n - i); xb = x[i:i + m]
        ok = False
        if s != 0 and np.isfinite(s):
            mant, ex = np.frexp(np.float64(s))          # |s| = mant * 2^ex, mant in [0.5, 1) -> binade exponent e = ex - 1
            e = int(ex) - 1
            if e - 23 > -120:
                u = np.float64(2.0) ** (e - 23)
                v = xb.astype(np.float64) / u               # exact (power of two scaling)
                if np.all(np.abs(v) < 2.0 ** 20):
                    r = np.rint(v)
                    tie = np.abs(v - np.trunc(v)) == 0.5
                    if not tie.any():
                        k0 = int(round(float(s) / u))
                        pre = k0 + np.cumsum(r.astype(np.int64))
                        sgn = 1 if k0 > 0 else -1
                        if np.all(sgn * pre > 2 ** 23) and np.all(sgn * pre < 2 ** 24):
                            s = np.float32(float(pre[-1]) * u); ok = True
        if ok: fast += 1
        else:
            for vv in xb: s = np.float32(s + vv)
            slow += 1
        i += m
    return s, fast, slow

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    for name, col in [("uniform x", synth.uniform(5, n, 3)[:, 0]), ("lidar x", synth.lidar(6, n)[:, 0]), ("lidar z", synth.lidar(6, n)[:, 2]),
                      ("lattice", synth.grid_ties(3, n, 3)[:, 1]), ("uniform-0.5", (synth.uniform(7, n, 3)[:, 0] - np.float32(0.5)))]:
        for blk in (256, 1024):
            a = seq_sum(col); b, f, sl = block_sum(col, blk)
            print(f"{name:12s} n={n} blk={blk}: equal={a.tobytes() == b.tobytes()} fast blocks {f} slow {sl} ({100 * sl / (f + sl):.1f}% sequential)")


# ---- tie-aware model, structured like the device code (32 lanes x EPL contiguous elements per tile) -------------------
def tile_sum_lanes(s, xb, EPL):
    """returns (ok, s_out).  Mirrors seq_sum_tile in csrc/seqsum.cuh."""
    bits = np.float32(s).view(np.uint32)
    ef = (int(bits) >> 23) & 0xff
    if ef < 24 or ef == 255: return False, s
    u = np.float32(2.0) ** np.float32(ef - 127 - 23)
    scale = np.float32(1.0) / u
    k_in = int(np.float32(s) * scale)
    M = len(xb); L = 32
    assert M == L * EPL
    v = (xb.astype(np.float32) * scale).astype(np.float32)
    if not np.all(np.abs(v) < 2.0 ** 20): return False, s
    r = np.rint(v).astype(np.int64)
    tie = np.abs(v - r.astype(np.float32)) == 0.5
    alt = r + np.where(v > r, 1, -1)
    sum0 = np.zeros(L, np.int64); mn0 = np.zeros(L, np.int64); mx0 = np.zeros(L, np.int64)
    has = np.zeros(L, bool); dlt = np.zeros(L, np.int64); pabs = np.zeros(L, np.int64)
    for l in range(L):
        q = 0; sm = 0; mn = 1 << 60; mx = -(1 << 60); seen = False
        for t in range(EPL):
            i = l * EPL + t
            if tie[i]:
                inc = alt[i] if q else r[i]
                if not seen:
                    seen = True
                    inc_other = r[i] if q else alt[i]          # what hypothesis p_in = 1 would add here
                    dlt[l] = inc_other - inc
                q = 0
            else:
                inc = r[i]; q ^= int(inc) & 1
            sm += inc; mn = min(mn, sm); mx = max(mx, sm)
        sum0[l] = sm; mn0[l] = mn; mx0[l] = mx; has[l] = seen; pabs[l] = q
    # parity entering each lane: p_out = has ? pabs : p_in ^ (sum0 & 1)
    p = k_in & 1; tot = 0; lo = 1 << 60; hi = -(1 << 60)
    for l in range(L):
        sl = sum0[l] + (dlt[l] if p else 0)
        lo = min(lo, tot + mn0[l] - 1); hi = max(hi, tot + mx0[l] + 1)    # the other hypothesis differs by at most 1
        tot += sl
        p = pabs[l] if has[l] else p ^ (int(sum0[l]) & 1)
        # note: without a tie the lane's own parity track started from 0: pabs is relative; with a tie it is absolute
    sg = 1 if k_in > 0 else -1
    if not (sg * (k_in + lo) > 2 ** 23 and sg * (k_in + hi) > 2 ** 23 and sg * (k_in + lo) < 2 ** 24 and sg * (k_in + hi) < 2 ** 24):
        return False, s
    return True, np.float32(np.float32(k_in + tot) * u)


def block_sum2(x, M=512):
    s = np.float32(0); i = 0; n = len(x); fast = slow = 0
    while i + M <= n:
        ok, s2 = tile_sum_lanes(s, x[i:i + M], M // 32) if s != 0 else (False, s)
        if ok: s = s2; fast += 1
        else:
            for vv in x[i:i + M]: s = np.float32(s + vv)
            slow += 1
        i += M
    for vv in x[i:]: s = np.float32(s + vv)
    return s, fast, slow


if __name__ == "__main__":
    print("---- tie-aware tiles ----")
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    g = np.random.default_rng(1)
    cols = [("uniform x", synth.uniform(5, n, 3)[:, 0]), ("lidar x", synth.lidar(6, n)[:, 0]), ("lidar y", synth.lidar(6, n)[:, 1]), ("lidar z", synth.lidar(6, n)[:, 2]),
            ("lattice", synth.grid_ties(3, n, 3)[:, 1]), ("uniform-0.5", (synth.uniform(7, n, 3)[:, 0] - np.float32(0.5))),
            ("halves", (g.integers(-50, 50, n) * 0.5).astype(np.float32)), ("gauss*30", (g.standard_normal(n) * 30).astype(np.float32)),
            ("sorted lidar x", np.sort(synth.lidar(6, n)[:, 0]))]
    for name, col in cols:
        for M in (256, 512):
            a = seq_sum(col); b, f, sl = block_sum2(col, M)
            print(f"{name:14s} n={n} tile={M}: equal={a.tobytes() == b.tobytes()} fast {f} slow {sl} ({100 * sl / max(f + sl, 1):.1f}% sequential)")
