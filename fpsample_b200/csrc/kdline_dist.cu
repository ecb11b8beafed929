// kdline_dist.cu -- QuickFPS kd-line SAMPLING for clouds beyond one SM: one thread-block cluster per cloud with the
// buckets (kd leaves) DISTRIBUTED over the CTAs -- CTA r owns the consecutive leaves [r*NB, (r+1)*NB), i.e. one
// subtree, spatially compact -- and several picks resolved per cluster-wide exchange.  Points (permuted SoA
// coordinates + running distances) stay in the per-cloud region in L2.
//
// Same observable result as the reference's lazy bucket scheme (src/_ext/KDNode.h:120-166,
// src/_ext/KDLineTree.h:56-85): exact FPS over the permuted array, ties to the lowest position.
// One iteration:
//   1. every CTA (warp 0, lane = owned bucket) extracts its M largest bucket maxima (key = distance bits, ~position)
//      and the next one as a bound for everything it does not send;
//   2. all-gather over distributed shared memory: each CTA stores its M candidates {key, second-largest distance
//      of the bucket, max point coordinates} + bound into every CTA of the cluster (st.shared::cluster) and
//      arrives on the receivers' mbarriers -- no cluster barrier, no global memory;
//   3. every CTA sorts the same 32 candidates (parallel rank computation) and accepts the longest prefix that is
//      provably the next J picks of the sequential recurrence: candidate j is above every bound, no earlier pick
//      of the batch lowers j's max point (dist(P_j, P_i) > val_j), and what is left of an earlier pick's bucket
//      stays below it (snd_i < val_j); running distances only decrease, so nothing else can overtake;
//   4. owners test the J picks against their buckets with the reference's own rules -- box bound
//      (KDNode.h:105-118) against the bucket max: drop; distance to the bucket's max point (KDNode.h:122-123):
//      defer to the pending list or flush;
//   5. flushes run on the owning CTA, all warps, 256 positions per work item: every pending sample is applied in
//      one pass (KDNode.h:147-161), bucket max / lowest position / second-largest distance re-derived.
// Every CTA computes the same J from the same exchanged data, so the cluster stays in lockstep without a barrier.
#include <cfloat>

#include "common.cuh"
#include "engine.h"

namespace fps {

constexpr u32 X_NC = 32;       // candidates per iteration over the whole cluster (one per lane of the sorting warp)
constexpr u32 X_RF = 12;       // a pending list this long flushes at the end of the iteration
constexpr u32 X_RCAP = X_RF + X_NC;   // an iteration appends at most X_NC samples
constexpr u32 X_WCH = 128;     // positions per flush work item (4 per lane)
constexpr u32 X_MAXITEMS = 1024;
#ifndef XDBG
#define XDBG 0
#endif
__device__ u64 g_dist_dbg[16];

struct DistArgs {
    unsigned char *region;
    size_t region_stride;
    const u64 *starts;
    u64 *out;
    u32 B, n, npad, dim, k, S, nlo_pad, NB, M, msh;
};

__device__ __forceinline__ void x_st_cluster_v4(u32 caddr, uint4 v) {
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(caddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 x_lds_v4(u32 a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}

// exchanged / partial entry: words {value bits, key low (0xfffffffe - position), second-largest bits, c[0..DIM)}
template <int DIM>
struct XEntry {
    static constexpr int W = (3 + DIM + 3) / 4 * 4;
    u32 w[W];
};

template <int DIM>
struct XTab {   // the iteration's candidates in descending key order
    u32 L0, bad, J, pad;
    u32 pos[X_NC];
    float val[X_NC];
    float snd[X_NC];
    float c[DIM][X_NC];
};

__device__ __forceinline__ void x_st_async_v4(u32 caddr, uint4 v, u32 cbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(caddr),
                 "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(cbar)
                 : "memory");
}
__device__ __forceinline__ void x_arrive_expect_tx_remote(u32 cbar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.relaxed.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void x_mbar_wait_cta(u32 bar, u32 parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "XWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra XWAIT_%=;\n\t}" ::"r"(bar),
        "r"(parity)
        : "memory");
}

template <int DIM>
__global__ void __launch_bounds__(512, 1) kdline_dist_kernel(DistArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using E = XEntry<DIM>;
    constexpr u32 EW = E::W;
    const u32 C = cluster_nctarank(), rank = cluster_ctarank();
    const u32 ncl = gridDim.x / C, cl = blockIdx.x / C;
    const u32 tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    const u32 T = blockDim.x, NW = T >> 5;
    const u32 npad = a.npad, dim = a.dim, NB = a.NB, M = a.M, msh = a.msh;
    const u32 NCAND = C * M;   // <= 32

    // ---- shared memory carve ---------------------------------------------------------------------------------------
    u64 *xbar = reinterpret_cast<u64 *>(smem_raw);                       // [2]
    u64 *kbuf = xbar + 2;                                                // [32] owner keys (top-M selection)
    u32 *rankv = reinterpret_cast<u32 *>(kbuf + 32);                     // [32]
    u32 *fl_b = rankv + 32;                                              // [32] flagged local bucket
    u32 *fl_n = fl_b + 32;                                               // [32] its pending count
    u32 *fl_lo = fl_n + 32, *fl_hi = fl_lo + 32, *fl_first = fl_hi + 32; // [32] each
    u32 *fl_span = fl_first + 32;                                        // [32] positions per work item
    u32 *fl_i0 = fl_span + 32;                                           // [33] first work item
    u32 *misc = fl_i0 + 36;                                              // [4]  nfl
    unsigned short *ptab = reinterpret_cast<unsigned short *>(misc + 4); // [496 -> 512] pair p -> i | j << 8
    XTab<DIM> *tab = reinterpret_cast<XTab<DIM> *>(ptab + 512);
    E *stage = reinterpret_cast<E *>(tab + 1);                           // [M + 1]
    E *xchg = stage + (M + 1);                                           // [2][C][M + 1]
    E *part = xchg + 2 * C * (M + 1);                                    // [X_MAXITEMS]
    float *pend = reinterpret_cast<float *>(part + X_MAXITEMS);          // [X_RCAP][DIM][32]

    if (tid == 0) {
        mbar_init(smem_u32(&xbar[0]), C);
        mbar_init(smem_u32(&xbar[1]), C);
        fence_mbar_init_cluster();
    }
    if (tid < 32) rankv[tid] = 0;
    for (u32 p = tid; p < 496; p += T) {
        u32 j = 1;
        while ((j + 1) * j / 2 <= p) ++j;
        ptab[p] = (unsigned short)((p - j * (j - 1) / 2) | (j << 8));
    }
    __syncthreads();
    cluster_sync_all();
    u32 xphase = 0;   // bit p: parity to wait for on xbar[p]
    u32 xpar = 0;

    for (u32 cloud = cl; cloud < a.B; cloud += ncl) {
        unsigned char *rg = a.region + (size_t)cloud * a.region_stride;
        const float *q = reinterpret_cast<const float *>(rg);
        float *dis = reinterpret_cast<float *>(rg) + (size_t)dim * npad;
        const u32 *perm = reinterpret_cast<const u32 *>(dis + npad);
        const u32 *nlo = perm + npad;
        const float *fbox = reinterpret_cast<const float *>(nlo + a.nlo_pad);
        u64 *out = a.out + (size_t)cloud * a.k;

        // ---- owner state (warp 0: lane l owns bucket rank*NB + l) -----------------------------------------------------
        bool valid = false;
        u32 blo = 0, bhi = 0, pos = 0, np = 0, first = 1;
        float mx = FLT_MAX, snd = 0.0f;
        float lo[DIM], hi[DIM], mc[DIM], clo[DIM], chi[DIM];
#pragma unroll
        for (int c = 0; c < DIM; ++c) {
            lo[c] = FLT_MAX;   // empty lanes do not widen the CTA box
            hi[c] = -FLT_MAX;
            mc[c] = 0.0f;
            clo[c] = chi[c] = 0.0f;
        }
        if (warp == 0) {
            if (lane < NB) {
                const u32 b = rank * NB + lane;
                blo = nlo[b];
                bhi = nlo[b + 1];
                if (bhi > blo) {
                    valid = true;
#pragma unroll
                    for (int c = 0; c < DIM; ++c) {
                        lo[c] = (c < (int)dim) ? fbox[(size_t)b * 2 * dim + c] : 0.0f;
                        hi[c] = (c < (int)dim) ? fbox[(size_t)b * 2 * dim + dim + c] : 0.0f;
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < DIM; ++c) {   // box of everything this CTA owns: one test screens a far-away pick
                clo[c] = ord2f(__reduce_min_sync(FULL, f2ord(lo[c])));
                chi[c] = ord2f(__reduce_max_sync(FULL, f2ord(hi[c])));
            }
        }
        // boot: the first sample = the point at POSITION start (wrapper.hpp:54-55); every bucket scans it
        // (KDNode::init) -- bucket maxima start at FLT_MAX, so the ordinary tests flush them all
        if (tid == 0) {
            const u32 cur = a.starts ? (u32)a.starts[cloud] : 0u;
            tab->L0 = 1;
            tab->bad = 0;
            tab->pos[0] = cur;
            for (u32 c = 0; c < DIM; ++c) tab->c[c][0] = c < dim ? __ldg(q + (size_t)c * npad + cur) : 0.0f;
        }
        __syncthreads();
        bool boot = true;
#if XDBG
        u64 dbg[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        const bool dbg_on = (tid == 0 && blockIdx.x == 0);
#endif

        for (u32 t = 0; t < a.k;) {
            u32 J = 1;
#if XDBG
            long long c0 = clock64(), c1 = c0, c2 = c0, c3 = c0;
#endif
            if (!boot) {
                // ---- 1. my CTA's M largest bucket maxima + the next one as the bound: rank of every owner key ------------
                if (warp == 0) {
                    const u64 mykey = valid ? make_key(mx, 0xfffffffeu - pos) : 0ull;
                    kbuf[lane] = mykey;
                    if (lane <= M) {   // zero entries: a CTA may own fewer than M non-empty buckets
                        u32 *e = stage[lane].w;
#pragma unroll
                        for (u32 w = 0; w < EW; ++w) e[w] = 0u;
                    }
                    __syncwarp();
                    u32 myr = 0;
#pragma unroll 8
                    for (u32 m = 0; m < 32; ++m) {
                        const u64 km = kbuf[m];
                        myr += ((km > mykey) | ((km == mykey) & (m < lane))) ? 1u : 0u;
                    }
                    if (myr <= M && mykey != 0ull) {
                        u32 *e = stage[myr].w;
                        e[0] = (u32)(mykey >> 32);
                        e[1] = (u32)mykey;
                        if (myr < M) {
                            e[2] = __float_as_uint(snd);
#pragma unroll
                            for (int c = 0; c < DIM; ++c) e[3 + c] = __float_as_uint(mc[c]);
                        }
                    }
                    __syncwarp();
                    // ---- 2. all-gather: lane c sends my entries to CTA c (st.async + complete_tx: no fence) ---------------
                    if (lane < C) {
                        const u32 src = smem_u32(stage);
                        const u32 dst = mapa(smem_u32(xchg + ((size_t)xpar * C + rank) * (M + 1)), lane);
                        const u32 bar = mapa(smem_u32(&xbar[xpar]), lane);
                        const u32 nch = (M + 1) * EW / 4;
                        x_arrive_expect_tx_remote(bar, 16u * nch);
#pragma unroll 1
                        for (u32 w = 0; w < nch; ++w) x_st_async_v4(dst + 16 * w, x_lds_v4(src + 16 * w), bar);
                    }
                }
#if XDBG
                c1 = clock64();
#endif
                x_mbar_wait_cta(smem_u32(&xbar[xpar]), (xphase >> xpar) & 1u);
#if XDBG
                c2 = clock64();
#endif
                xphase ^= 1u << xpar;
                const E *xc = xchg + (size_t)xpar * C * (M + 1);
                xpar ^= 1u;
                // ---- 3. sort the C*M candidates: thread (i, slice) counts the candidates of its slice in front of i ------
                {
                    const u32 i = tid & 31u;
                    if (i < NCAND) {
                        const u32 *ei = xc[(i >> msh) * (M + 1) + (i & (M - 1))].w;
                        const u64 ki = ((u64)ei[0] << 32) | ei[1];
                        u32 cnt = 0;
#pragma unroll 1
                        for (u32 m = warp; m < NCAND; m += NW) {
                            const u32 *em = xc[(m >> msh) * (M + 1) + (m & (M - 1))].w;
                            const u64 km = ((u64)em[0] << 32) | em[1];
                            cnt += ((km > ki) | ((km == ki) & (m < i))) ? 1u : 0u;
                        }
                        if (cnt) atomicAdd(&rankv[i], cnt);
                    }
                }
                __syncthreads();
                if (warp == 0) {
                    const bool in = lane < NCAND;
                    const u32 *ei = xc[in ? (lane >> msh) * (M + 1) + (lane & (M - 1)) : 0].w;
                    const u32 khi = in ? ei[0] : 0u, klo = in ? ei[1] : 0u;
                    // bound = largest of the CTAs' next keys: value first, then the low word among equal values
                    u32 bhi_ = 0, blo_ = 0;
                    if (lane < C) {
                        const u32 *eb = xc[lane * (M + 1) + M].w;
                        bhi_ = eb[0];
                        blo_ = eb[1];
                    }
                    const u32 Bhi = __reduce_max_sync(FULL, bhi_);
                    const u32 Blo = __reduce_max_sync(FULL, bhi_ == Bhi ? blo_ : 0u);
                    const u32 r = in ? rankv[lane] : 31u;
                    if (in) rankv[lane] = 0;
                    const bool nz = (khi | klo) != 0u;
                    const bool elig = in && nz && (khi > Bhi || (khi == Bhi && klo > Blo));
                    const u32 inel = __reduce_or_sync(FULL, (in && !elig) ? (1u << r) : 0u);
                    if (in) {
                        tab->pos[r] = 0xfffffffeu - klo;
                        tab->val[r] = __uint_as_float(khi);
                        tab->snd[r] = __uint_as_float(ei[2]);
#pragma unroll
                        for (int c = 0; c < DIM; ++c) tab->c[c][r] = __uint_as_float(ei[3 + c]);
                    }
                    if (lane == 0) {
                        tab->L0 = inel ? (u32)__ffs(inel) - 1u : NCAND;
                        tab->bad = 0u;
                    }
                }
                __syncthreads();
                const u32 L0 = tab->L0;
                if (L0 > 1) {   // pair checks, one pair (i < j) per thread
                    const u32 NPAIR = L0 * (L0 - 1) / 2;
#pragma unroll 1
                    for (u32 p = tid; p < NPAIR; p += T) {
                        const u32 ij = ptab[p];
                        const u32 i = ij & 255u, j = ij >> 8;
                        float Pi[DIM], Pj[DIM];
#pragma unroll
                        for (int c = 0; c < DIM; ++c) {
                            Pi[c] = tab->c[c][i];
                            Pj[c] = tab->c[c][j];
                        }
                        const float vj = tab->val[j];
                        if (!(sqdist<DIM>(Pj, Pi) > vj) || !(tab->snd[i] < vj)) atomicOr(&tab->bad, 1u << j);
                    }
                    __syncthreads();   // L0 is uniform, so is this barrier
                }
                const u32 badm = tab->bad;
                J = badm ? (u32)__ffs(badm) - 1u : X_NC;
                if (J > L0) J = L0;
                if (J > a.k - t) J = a.k - t;
            }
#if XDBG
            c3 = clock64();
#endif
            // ---- 4. owners: the J accepted picks against my buckets -------------------------------------------------------
            u32 myidx = 0;
            bool flush = false;
            if (warp == 0) {
                const float cthr = __uint_as_float(__reduce_max_sync(FULL, valid ? __float_as_uint(mx) : 0u));
                bool relv = false;
                if (lane < J) {
                    float pj[DIM];
#pragma unroll
                    for (int c = 0; c < DIM; ++c) pj[c] = tab->c[c][lane];
                    relv = boxdist<DIM>(pj, clo, chi) < cthr;
                }
                u32 rel = __ballot_sync(FULL, relv);
                bool hitany = false;
#pragma unroll 1
                while (rel) {
                    const u32 j = __ffs(rel) - 1;
                    rel &= rel - 1;
                    float pc[DIM];
#pragma unroll
                    for (int c = 0; c < DIM; ++c) pc[c] = tab->c[c][j];
                    const bool touch = boxdist<DIM>(pc, lo, hi) < mx;          // KDNode.h:126-130
                    const bool hit = !(sqdist<DIM>(mc, pc) > mx);              // KDNode.h:122-123
                    if (valid && (touch || hit)) {
#pragma unroll
                        for (int c = 0; c < DIM; ++c) pend[((size_t)np * DIM + c) * 32 + lane] = pc[c];
                        ++np;
                        hitany |= hit;
                    }
                }
                flush = valid && np > 0 && (hitany || np >= X_RF);
                const u32 fm = __ballot_sync(FULL, flush);
                // work items: X_WCH positions each; a huge bucket gets longer items so that it fits the partial slots
                u32 span = X_WCH;
                if (flush && (bhi - blo) > X_WCH * X_MAXITEMS) span = (((bhi - blo) + X_MAXITEMS - 1) / X_MAXITEMS + X_WCH - 1) / X_WCH * X_WCH;
                const u32 items = flush ? (bhi - blo + span - 1) / span : 0u;
                u32 inc = items;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const u32 y = __shfl_up_sync(FULL, inc, o);
                    if ((int)lane >= o) inc += y;
                }
                myidx = __popc(fm & ((1u << lane) - 1u));
                if (flush) {
                    fl_b[myidx] = lane;
                    fl_n[myidx] = np;
                    fl_lo[myidx] = blo;
                    fl_hi[myidx] = bhi;
                    fl_first[myidx] = first;
                    fl_span[myidx] = span;
                    fl_i0[myidx] = inc - items;
                }
                if (lane == 31) {
                    misc[0] = __popc(fm);
                    fl_i0[__popc(fm)] = inc;
                }
            }
            if (rank == 0 && tid < J) out[t + tid] = tab->pos[tid];   // positions now, original ids at the end
            __syncthreads();
#if XDBG
            const long long c4 = clock64();
#endif
            // ---- 5. flush work items, any warp; rounds of whole buckets only if the items exceed the partial slots ----------
            const u32 nfl = misc[0];
#pragma unroll 1
            for (u32 f0 = 0; f0 < nfl;) {
                u32 f1 = f0 + 1;
                const u32 ib = fl_i0[f0];
                while (f1 < nfl && fl_i0[f1 + 1] - ib <= X_MAXITEMS) ++f1;
                const u32 ie = fl_i0[f1];
#pragma unroll 1
                for (u32 item = ib + warp; item < ie; item += NW) {
                    u32 f = f0;
                    while (f + 1 < f1 && fl_i0[f + 1] <= item) ++f;
                    const u32 span = fl_span[f];
                    const u32 p0 = fl_lo[f] + (item - fl_i0[f]) * span, p1 = min(fl_hi[f], p0 + span);
                    const u32 nref = fl_n[f], lb = fl_b[f], frst = fl_first[f];
                    float best = -1.0f, sec = 0.0f;
                    u32 bi = 0;
                    float bc[DIM];
#pragma unroll
                    for (int c = 0; c < DIM; ++c) bc[c] = 0.0f;
#pragma unroll 1
                    for (u32 pp = p0; pp < p1; pp += X_WCH) {
                        float x[DIM][X_WCH / 32], v[X_WCH / 32];
#pragma unroll
                        for (int u = 0; u < (int)(X_WCH / 32); ++u) {
                            const u32 p = pp + u * 32 + lane;
                            const bool in = p < p1;
#pragma unroll
                            for (int c = 0; c < DIM; ++c) x[c][u] = (in && c < (int)dim) ? __ldg(q + (size_t)c * npad + p) : 0.0f;
                            v[u] = (in && !frst) ? __ldcg(dis + p) : FLT_MAX;   // Point.h:61-65
                        }
#pragma unroll 1
                        for (u32 r = 0; r < nref; ++r) {
                            float ref[DIM];
#pragma unroll
                            for (int c = 0; c < DIM; ++c) ref[c] = pend[((size_t)r * DIM + c) * 32 + lb];
#pragma unroll
                            for (int u = 0; u < (int)(X_WCH / 32); ++u) {
                                float pt[DIM];
#pragma unroll
                                for (int c = 0; c < DIM; ++c) pt[c] = x[c][u];
                                v[u] = fminf(v[u], sqdist<DIM>(pt, ref));   // std::min(dis, d), Point.h:82-86
                            }
                        }
#pragma unroll
                        for (int u = 0; u < (int)(X_WCH / 32); ++u) {
                            const u32 p = pp + u * 32 + lane;
                            if (p < p1) {
                                __stcg(dis + p, v[u]);
                                if (v[u] > best) {   // ascending p: the first maximum = lowest position stays
                                    sec = best < 0.0f ? 0.0f : best;
                                    best = v[u];
                                    bi = p;
#pragma unroll
                                    for (int c = 0; c < DIM; ++c) bc[c] = x[c][u];
                                } else {
                                    sec = fmaxf(sec, v[u]);
                                }
                            }
                        }
                    }
                    const u64 key = best < 0.0f ? 0ull : make_key(best, 0xfffffffeu - bi);
                    const u64 wk = warp_max_key(key);
                    const bool iwin = (key == wk) && key != 0ull;
                    const u32 SND = __reduce_max_sync(FULL, __float_as_uint(iwin ? sec : fmaxf(best, 0.0f)));
                    if (iwin) {
                        u32 *e = part[item - ib].w;
                        e[0] = (u32)(wk >> 32);
                        e[1] = (u32)wk;
                        e[2] = SND;
#pragma unroll
                        for (int c = 0; c < DIM; ++c) e[3 + c] = __float_as_uint(bc[c]);
                    }
                }
                __syncthreads();
                // ---- 6. owners take the results: a one-item bucket directly, the others merged by the whole warp ---------------
                if (warp == 0) {
                    const bool mine = flush && myidx >= f0 && myidx < f1;
                    const u32 mi0 = mine ? fl_i0[myidx] : 0u, mi1 = mine ? fl_i0[myidx + 1] : 0u;
                    if (mine && mi1 - mi0 == 1) {
                        const u32 *e = part[mi0 - ib].w;
                        mx = __uint_as_float(e[0]);
                        pos = 0xfffffffeu - e[1];
                        snd = __uint_as_float(e[2]);
#pragma unroll
                        for (int c = 0; c < DIM; ++c) mc[c] = __uint_as_float(e[3 + c]);
                        np = 0;
                        first = 0;
                    }
                    u32 multi = __ballot_sync(FULL, mine && mi1 - mi0 > 1);
#pragma unroll 1
                    while (multi) {
                        const u32 ol = __ffs(multi) - 1;   // owner lane of a bucket with several items
                        multi &= multi - 1;
                        const u32 i0 = __shfl_sync(FULL, mi0, ol), i1 = __shfl_sync(FULL, mi1, ol);
                        u64 key = 0;
                        u32 sn = 0, wi = 0;
#pragma unroll 1
                        for (u32 i = i0 + lane; i < i1; i += 32) {   // a lane folds its items: keep the best, bound the rest
                            const u32 *e = part[i - ib].w;
                            const u64 ke = ((u64)e[0] << 32) | e[1];
                            if (ke > key) {
                                sn = max(max(sn, (u32)(key >> 32)), e[2]);
                                key = ke;
                                wi = i;
                            } else {
                                sn = max(sn, e[0]);
                            }
                        }
                        const u64 K = warp_max_key(key);
                        const bool iw = key == K && key != 0ull;
                        const u32 SN = __reduce_max_sync(FULL, iw ? sn : (u32)(key >> 32));
                        const u32 wsrc = __ffs(__ballot_sync(FULL, iw)) - 1;
                        const u32 witem = __shfl_sync(FULL, wi, wsrc);
                        if (lane == ol) {
                            const u32 *e = part[witem - ib].w;
                            mx = __uint_as_float((u32)(K >> 32));
                            pos = 0xfffffffeu - (u32)K;
                            snd = __uint_as_float(SN);
#pragma unroll
                            for (int c = 0; c < DIM; ++c) mc[c] = __uint_as_float(e[3 + c]);
                            np = 0;
                            first = 0;
                        }
                    }
                }
                f0 = f1;
                if (f0 < nfl) __syncthreads();   // the partial slots are reused by the next round
            }
            t += J;
            boot = false;
#if XDBG
            if (dbg_on) {
                const long long c5 = clock64();
                dbg[0] += 1;
                dbg[1] += J;
                dbg[2] += (u64)(c1 - c0);   // candidates + send
                dbg[3] += (u64)(c2 - c1);   // wait for the cluster
                dbg[4] += (u64)(c3 - c2);   // sort + pair checks
                dbg[5] += (u64)(c4 - c3);   // tests
                dbg[6] += (u64)(c5 - c4);   // flush + merge
                dbg[7] += nfl;              // flushed buckets of this CTA
                dbg[8] += nfl ? fl_i0[nfl] : 0;   // items
            }
#endif
        }
#if XDBG
        if (dbg_on)
            for (int i = 0; i < 10; ++i) g_dist_dbg[i] = dbg[i];
#endif
        // ---- positions -> original ids (wrapper.hpp:57-59) ------------------------------------------------------------------
        __threadfence_block();
        __syncthreads();
        if (rank == 0) {
            for (u32 t = tid; t < a.k; t += T) {
                const u32 p = (u32)__ldcg(reinterpret_cast<const unsigned long long *>(out + t));
                out[t] = perm[p];
            }
        }
        __syncthreads();
    }
    cluster_sync_all();  // nobody leaves while a peer may still write into its shared memory
}

// ======================================================================================================
//  host side
// ======================================================================================================
static int pad_dim_x(int dim) { return dim <= 2 ? 2 : dim == 3 ? 3 : dim == 4 ? 4 : dim <= 6 ? 6 : 8; }

template <int DIM>
static size_t dist_smem(size_t C, size_t M) {
    size_t b = 2 * 8 + 32 * 8 + (32 * 7 + 36 + 4) * 4 + 512 * 2 + sizeof(XTab<DIM>) + sizeof(XEntry<DIM>) * ((M + 1) + 2 * C * (M + 1) + X_MAXITEMS) +
               (size_t)X_RCAP * DIM * 32 * 4;
    return b + 64;
}
static size_t dist_smem_dim(int dimp, size_t C, size_t M) {
    switch (dimp) {
        case 2: return dist_smem<2>(C, M);
        case 3: return dist_smem<3>(C, M);
        case 4: return dist_smem<4>(C, M);
        case 6: return dist_smem<6>(C, M);
        default: return dist_smem<8>(C, M);
    }
}

bool plan_kdline_dist(size_t n, size_t dim, size_t h, size_t B, int n_sms, DistPlan *pl) {
    if (dim == 0 || dim > 8 || h == 0 || h > 9 || n == 0 || B == 0) return false;
    // EXPERIMENTAL, opt-in (FPS_B200_DIST=1): parity-green, but the lockstep iteration (exchange -> sort -> tests ->
    // flush, each a chain of dependent reductions / barriers) measured slower per pick than the asynchronous
    // coordinator/worker kernel on every BASELINE config (cfg4: 1370 vs 826 cycles per pick); see DESIGN.md.
    const char *en = getenv("FPS_B200_DIST");
    if (!en || atoi(en) == 0) return false;
    const size_t S = (size_t)1 << h;
    // each CTA owns at most 32 buckets (one per lane of its owner warp)
    u32 C = 1;
    while (S / C > 32) C *= 2;
    if (C > 16) return false;
    // few clouds: spread each over more SMs (fewer buckets, hence fewer flushes, per CTA and per iteration)
    while (C < 16 && S / (C * 2) >= 4 && B * C * 2 <= (size_t)n_sms && n / (C * 2) >= 4096) C *= 2;
    if (const char *e = getenv("FPS_B200_DIST_C")) {
        u32 c = (u32)atoi(e);
        if (c >= C && c <= 16 && (c & (c - 1)) == 0 && S / c >= 1) C = c;
    }
    const u32 NB = (u32)(S / C);
    u32 M = X_NC / C;
    if (M > NB) M = NB;
    if (M > 16) M = 16;
    if (M < 1) return false;
    u32 threads = n / S >= 1024 ? 512 : 256;
    if (const char *e = getenv("FPS_B200_DIST_T")) threads = (u32)atoi(e) >= 512 ? 512 : 256;
    pl->dimp = pad_dim_x((int)dim);
    pl->C = C;
    pl->NB = NB;
    pl->M = M;
    pl->threads = threads;
    pl->smem = dist_smem_dim(pl->dimp, C, M);
    if (pl->smem > 200 * 1024) return false;
    size_t clusters = (size_t)n_sms / C;
    if (clusters > B) clusters = B;
    if (clusters < 1) clusters = 1;
    pl->clusters = (u32)clusters;
    return true;
}

template <int DIM>
static cudaError_t launch_dist_t(const DistPlan &pl, const DistArgs &a, cudaStream_t st) {
    auto kern = kdline_dist_kernel<DIM>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) return e;
    if (pl.C > 8) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(pl.clusters * pl.C);
    cfg.blockDim = dim3(pl.threads);
    cfg.dynamicSmemBytes = pl.smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = pl.C;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, a);
}

cudaError_t dist_debug_counters(u64 *out16) { return cudaMemcpyFromSymbol(out16, g_dist_dbg, sizeof(u64) * 16); }

cudaError_t launch_kdline_dist(const DistPlan &pl, unsigned char *region, size_t region_stride, const u64 *starts, u64 *out,
                               u32 B, u32 n, u32 dim, u32 k, u32 h, cudaStream_t st) {
    DistArgs a;
    a.region = region;
    a.region_stride = region_stride;
    a.starts = starts;
    a.out = out;
    a.B = B;
    a.n = n;
    a.npad = (n + 31) & ~31u;
    a.dim = dim;
    a.k = k;
    a.S = 1u << h;
    a.nlo_pad = (a.S + 1 + 31) & ~31u;
    a.NB = pl.NB;
    a.M = pl.M;
    a.msh = 0;
    while ((1u << a.msh) < pl.M) ++a.msh;
    cudaError_t e;
    switch (pl.dimp) {
        case 2: e = launch_dist_t<2>(pl, a, st); break;
        case 3: e = launch_dist_t<3>(pl, a, st); break;
        case 4: e = launch_dist_t<4>(pl, a, st); break;
        case 6: e = launch_dist_t<6>(pl, a, st); break;
        default: e = launch_dist_t<8>(pl, a, st); break;
    }
    count_launch();
    return e;
}

}  // namespace fps
