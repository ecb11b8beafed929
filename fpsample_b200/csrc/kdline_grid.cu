// kdline_grid.cu -- QuickFPS kd-line SAMPLING for ONE HUGE cloud (BASELINE.json cfg 4: 2^20 points -> 65536): the
// whole GPU works on the cloud, every point stays in SHARED MEMORY for all K picks, and the K-1 dependent picks are
// resolved in BATCHES: one grid-wide exchange decides the next J picks (J ~ 45 on average, up to 256) instead of one.
//
// Semantics (reference src/_ext/KDLineTree.h:56-85, src/_ext/KDNode.h:84-166, src/wrapper.hpp:54-59):
// exact FPS over the array the kd build permuted (kdbuild.cu wrote it into the cloud's region), started at POSITION
// start, running distance initialised to FLT_MAX (src/_ext/Point.h:61-65), ties to the lowest position.
//
// Layout: the permuted array is cut into SLICES of SL = 32 * PPT consecutive positions that never straddle a kd leaf
// (a slice's box is then as tight as its leaf's); slice g belongs to warp g % 32 of CTA g / 32; lane l holds the slice's
// positions (u * 32 + l) * 4 + e (u < PPT/4, e < 4): coordinates and running distances sit in shared memory as SoA
// float4, one conflict-free 128-bit access per 4 points.  A slice has a tight box and an exact maximum.
//
// One round (scripts/sim_grid.py is the CPU model of this protocol, bit-exact against the oracle at full size):
//   A  every warp applies the picks of the previous round that can lower something in its slice (the reference's own
//      point-to-box bound, KDNode.h:105-118, against the slice maximum: skipped work never changes a distance) and, if
//      touched, re-selects its 2 largest keys + the third as the slice BOUND.  key = distance bits << 32 | ~position.
//   B  every warp ranks its two keys among the CTA's 64; ranks 0..8 are PUBLISHED together with the largest slice bound:
//      ten 8-byte {distance, position | stamp} words per CTA in global memory, the stamp being the round number.
//   C  every CTA gathers every CTA's words (spinning on the stamps: the only grid-wide synchronisation, one L2 round
//      trip, no atomics, no fences) and selects -- redundantly, identically -- the candidates above EVERY CTA's bound
//      (every point a CTA did not publish sorts at or below its bound), sorts them (rank by counting), and accepts the
//      longest prefix in which no candidate is lowered by an earlier one: dist(P_j, P_i) >= val_j for all i < j.
//      Running distances only decrease, so that prefix IS the next J picks of the sequential recurrence, in order.
//
// What bounds a round is the number of DEPENDENT instructions on one warp's path (~5 cycles each; measured with
// scripts/micro/sel.cu, icache.cu, xchg.cu), not arithmetic: every phase below is written to keep that chain short.
#include <cfloat>
#include <cstdio>

#include "common.cuh"
#include "engine.h"

namespace fps {

constexpr u32 G_T = 1024;        // threads per CTA
constexpr u32 G_W = 32;          // warps per CTA
constexpr u32 G_M = 8;           // published candidates per CTA (+ the 9th key as part of the bound)
constexpr u32 G_NK = G_M + 2;    // 8-byte words a CTA publishes: keys 0..8 and its largest slice bound
constexpr u32 G_ECAP = 256;      // eligible candidates per round at most
constexpr u32 G_MAXG = 160;      // CTAs at most
constexpr u32 G_NONE = 0xffffffffu;
constexpr u32 G_LOW = 0xfffffffeu;   // key low word = G_LOW - position
constexpr u32 G_PBITS = 21;          // published position field; the stamp lives above it
constexpr u32 G_PNONE = (1u << G_PBITS) - 1u;   // "no key"

#ifndef GDBG
#define GDBG 0   // 1: per-phase clock64 counters of CTA 0 (scripts/run_one.py prints them)
#endif
#if GDBG
#define GCLK() clock64()
#else
#define GCLK() 0ll
#endif
__device__ u64 g_grid_dbg[16];
#if GDBG
__device__ u32 g_grid_trace[4096 * 160];   // [round][cta]: cycles of phase A (incl. the wait for the CTA's slowest warp)
#endif

struct GridArgs {
    unsigned char *region;
    size_t region_stride;
    const u64 *starts;
    u64 *out;
    uint4 *pub;          // [2][G][G_NK / 2] stamped 16-byte chunks (two 8-byte words each), zeroed before the launch
    u32 B, n, npad, dim, k, ppt, ecap, S, gc;
    const float *qv;     // IDS: per cloud [dim][npad] coordinates by virtual position (the input reversed, SoA)
    size_t qv_stride;    // floats
    u32 n_starts;        // IDS: forced first picks per cloud (original indices), >= 1
    u64 negzero;         // two binary32 -0.0 as an operand the compiler cannot see through (packed products, common.cuh)
};

// gpu-scope relaxed accesses: served by L2, never by a stale L1 line
__device__ __forceinline__ uint4 ldg_relaxed_v4(const uint4 *p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stg_relaxed_v2(void *p, u32 x, u32 y) {
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(x), "r"(y) : "memory");
}

// point -> box squared distance (KDNode.h:105-118) without branches
template <int DIM>
__device__ __forceinline__ float g_boxdist(const float (&r)[DIM], const float (&lo)[DIM], const float (&hi)[DIM]) {
    float acc = 0.0f;
#pragma unroll
    for (int j = 0; j < DIM; ++j) {
        const float e = fmaxf(fmaxf(__fsub_rn(r[j], hi[j]), __fsub_rn(lo[j], r[j])), 0.0f);
        const float e2 = __fmul_rn(e, e);
        acc = (j == 0) ? e2 : __fadd_rn(acc, e2);
    }
    return acc;
}

__device__ __forceinline__ void bar_sync_named(u32 id, u32 nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ float f4get(const float4 &v, int e) { return e == 0 ? v.x : e == 1 ? v.y : e == 2 ? v.z : v.w; }
__device__ __forceinline__ u64 key_max(u64 a, u64 b) { return a > b ? a : b; }

// FLAT = false: one huge cloud on the whole grid, every CTA merges its warps' keys and publishes its 8 largest.
// FLAT = true : a batch of medium clouds, GROUPS of gc <= 16 CTAs per cloud; every warp publishes its own 3 largest keys
//               + a bound (no CTA-level merge: with few CTAs per cloud the candidates per round would be too few).
// IDS = true  : vanilla FPS (src/lib.cpp:111-246) through the same machinery: the kd permutation only decides which points
//               share a slice; every point carries a VIRTUAL position n - 1 - original index, so that the lowest virtual
//               position among equal distances is the HIGHEST original index (the '>=' at lib.cpp:226); coordinates of a
//               pick are fetched from the reversed input (a.qv), starts are original indices, distances start at +inf.
template <int DIM, bool FLAT, bool IDS>
__global__ void __launch_bounds__(G_T, 1) kdline_grid_kernel(GridArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const u32 tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    const u32 G = a.gc, cta = blockIdx.x % a.gc, grp = blockIdx.x / a.gc, ngrp = gridDim.x / a.gc;   // CTAs per cloud, my rank, my group
    const u32 npad = a.npad, dim = a.dim, k = a.k, S = a.S;
    const u32 NU = a.ppt >> 2;              // float4 per lane per component
    const u32 SL = 32u * a.ppt;             // positions per slice (one warp)
    const u32 PCQ = G_W * NU * 32u;         // float4 per component per CTA
    const u32 ECAP = a.ecap;

    // ---- shared memory carve ----------------------------------------------------------------------------------
    float4 *pts = reinterpret_cast<float4 *>(smem_raw);                         // [DIM + 1][PCQ]
    u64 *gk = reinterpret_cast<u64 *>(pts + (size_t)(DIM + 1) * PCQ);           // [G][G_NK] gathered keys
    u64 *wtop = gk + (FLAT ? (size_t)G * G_W * 4 : (size_t)G_MAXG * G_NK);                                     // [G_W][4]: 2 keys, bound, pad
    u64 *ekey = wtop + G_W * 4;                                                 // [2][G_ECAP]
    u64 *red = ekey + 2 * G_ECAP;                                               // [8][4]
    u32 *tpos = reinterpret_cast<u32 *>(red + 32);                              // [G_ECAP]
    float *tval = reinterpret_cast<float *>(tpos + G_ECAP);                     // [G_ECAP]
    float *tc = tval + G_ECAP;                                                  // [DIM][G_ECAP]
    u32 *rel = reinterpret_cast<u32 *>(tc + (size_t)DIM * G_ECAP);              // [G_ECAP]
    int *cbox = reinterpret_cast<int *>(rel + G_ECAP);                          // [2 * DIM] ordered ints
    u32 *misc = reinterpret_cast<u32 *>(cbox + 2 * DIM);                        // [16]
    u32 *wflag = misc + 16;                                                     // [16] conflict flag per window
    u32 *lw = wflag + 16;                                                       // [2][8] lowered-candidate bit words
    u32 *pickw = lw + 16;                                                       // [8] pick mask words, [8] their exclusive prefix counts
    u32 *rowany = pickw + 16;                                                      // [G_ECAP]
    u32 *conf = rowany + G_ECAP;                                                // [G_ECAP][8] conflict bits
    u32 *slt = conf + G_ECAP * 8;                                               // [2][512] flat mode: first position / size of every slice
    u32 *cum = slt + 1024;                                                      // [S + 1] (see below), then IDS: [PCQ] uint4 virtual positions
    uint4 *pid = reinterpret_cast<uint4 *>((reinterpret_cast<uintptr_t>(cum + S + 1) + 15) & ~(uintptr_t)15);
    //                                             // [S + 1] first slice of every leaf
    enum { M_NREL = 0, M_STOP = 1, M_J = 2, M_E0 = 4 };

    float4 *pv = pts + (size_t)DIM * PCQ;   // running distances; padding slots hold -1 (never a candidate)

    u32 round = 0;                          // runs across clouds
#if GDBG
    u64 dbg[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const bool dbg_on = (tid == 0 && blockIdx.x == 0);
#endif

    constexpr u32 WK = FLAT ? 4u : 0u;                                   // flat: words a warp publishes (3 keys + bound)
    uint4 *pubg = a.pub + (size_t)grp * 2 * (FLAT ? G * G_W * 2 : G * (G_NK / 2));   // this group's two exchange buffers
    for (u32 cloud = grp; cloud < a.B; cloud += ngrp) {
        unsigned char *rg = a.region + (size_t)cloud * a.region_stride;
        const float *q = reinterpret_cast<const float *>(rg);
        const u32 *perm = reinterpret_cast<const u32 *>(rg) + (size_t)(dim + 1) * npad;
        const u32 *nlo = perm + npad;
        u64 *out = a.out + (size_t)cloud * k;
        const float *qf = IDS ? a.qv + (size_t)cloud * a.qv_stride : q;   // coordinates by (virtual) position

        if (tid < 2 * DIM) cbox[tid] = tid < DIM ? 0x7fffffff : (int)0x80000000;
        u32 wpos0 = 0, wcnt = 0;   // this warp's slice: positions [wpos0, wpos0 + wcnt) of the permuted array
        if constexpr (FLAT) {
            // ---- a slice = the largest kd SUBTREE (aligned block of 2^j leaves) that fits SL positions: spatially as
            //      compact as the tree makes it; a leaf beyond SL positions is cut into several slices.  If the cloud needs
            //      more slices than the group has warps, fall back to plain position ranges (any slicing is exact, the
            //      boxes are just looser).
            for (u32 b = tid; b <= S; b += G_T) cum[b] = __ldg(nlo + b);
            __syncthreads();
            if (tid == 0) {
                const u32 cap = G * G_W;
                u32 ns = 0, pos = 0;
                bool ok = true;
                while (pos < S && ok) {
                    u32 j = pos ? (u32)__ffs(pos) - 1u : 31u;
                    while ((1u << j) > S - pos) --j;
                    while (j > 0 && cum[pos + (1u << j)] - cum[pos] > SL) --j;
                    const u32 lo = cum[pos], cnt = cum[pos + (1u << j)] - lo;
                    const u32 need = (cnt + SL - 1) / SL;
                    if (ns + need > cap) ok = false;
                    else
                        for (u32 x = 0; x < need; ++x) {
                            slt[ns] = lo + x * SL;
                            slt[512 + ns] = min(SL, cnt - x * SL);
                            ++ns;
                        }
                    pos += 1u << j;
                }
                if (!ok) {
                    ns = 0;
                    for (u32 p0 = 0; p0 < a.n; p0 += SL, ++ns) {
                        slt[ns] = p0;
                        slt[512 + ns] = min(SL, a.n - p0);
                    }
                }
                for (; ns < cap; ++ns) slt[ns] = slt[512 + ns] = 0;
            }
            __syncthreads();
            wpos0 = slt[cta * G_W + warp];
            wcnt = slt[512 + cta * G_W + warp];
        } else {
        // ---- leaf b owns the slices [cum[b], cum[b+1]), ceil(size / SL) of them.  Warp 0 scans the leaf sizes. ----------
        if (warp == 0) {
            u32 carry = 0;
            for (u32 b0 = 0; b0 < S; b0 += 32) {
                const u32 b = b0 + lane;
                u32 x = b < S ? (__ldg(nlo + b + 1) - __ldg(nlo + b) + SL - 1) / SL : 0u;
                const u32 mine = x;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const u32 y = __shfl_up_sync(FULL, x, o);
                    if (lane >= (u32)o) x += y;
                }
                if (b < S) cum[b] = carry + x - mine;
                carry += __shfl_sync(FULL, x, 31);
            }
            if (lane == 0) cum[S] = carry;
        }
        __syncthreads();
        {
            const u32 g = cta * G_W + warp;
            if (g < cum[S]) {
                u32 lo = 0, hi = S;   // largest b with cum[b] <= g (empty leaves share their value with the next leaf)
                while (hi - lo > 1) {
                    const u32 mid = (lo + hi) >> 1;
                    if (cum[mid] <= g) lo = mid;
                    else hi = mid;
                }
                const u32 l0 = __ldg(nlo + lo), l1 = __ldg(nlo + lo + 1);
                wpos0 = l0 + (g - cum[lo]) * SL;
                wcnt = min(SL, l1 - wpos0);
            }
        }
        }
        float wlo[DIM], whi[DIM];
        {
            float mn[DIM], mx[DIM];
#pragma unroll
            for (int c = 0; c < DIM; ++c) {
                mn[c] = __int_as_float(0x7f800000);
                mx[c] = __int_as_float(0xff800000);
            }
            for (u32 u = 0; u < NU; ++u) {
                const u32 slot = (warp * NU + u) * 32u + lane;
                const u32 l0 = (u * 32u + lane) * 4u;
                float vv[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) vv[e] = (l0 + e < wcnt) ? (IDS ? __int_as_float(0x7f800000) : FLT_MAX) : -1.0f;
                if constexpr (IDS) {
                    u32 id[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) id[e] = (l0 + e < wcnt) ? a.n - 1u - __ldg(perm + wpos0 + l0 + e) : 0u;
                    pid[slot] = make_uint4(id[0], id[1], id[2], id[3]);
                }
                pv[slot] = make_float4(vv[0], vv[1], vv[2], vv[3]);
#pragma unroll
                for (int c = 0; c < DIM; ++c) {
                    float x[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        x[e] = 0.0f;
                        if (c < (int)dim && l0 + e < wcnt) {
                            x[e] = __ldg(q + (size_t)c * npad + wpos0 + l0 + e);
                            mn[c] = fminf(mn[c], x[e]);
                            mx[c] = fmaxf(mx[c], x[e]);
                        }
                    }
                    pts[(size_t)c * PCQ + slot] = make_float4(x[0], x[1], x[2], x[3]);
                }
            }
#pragma unroll
            for (int c = 0; c < DIM; ++c) {
                const int lo_o = __reduce_min_sync(FULL, f2ord(mn[c])), hi_o = __reduce_max_sync(FULL, f2ord(mx[c]));
                wlo[c] = ord2f(lo_o);
                whi[c] = ord2f(hi_o);
                if (c >= (int)dim) wlo[c] = whi[c] = 0.0f;
                if (lane == 0 && c < (int)dim && wcnt) {
                    atomicMin(&cbox[c], lo_o);
                    atomicMax(&cbox[DIM + c], hi_o);
                }
            }
        }
        float smax = 0.0f;   // exact maximum of the slice (warp-uniform)

        // The warp's 2 largest keys + the third (its bound).  One pass: every lane keeps its three best (ascending
        // positions, strict '>': the lowest position wins a tie), the warp then pops three times.
        auto reselect = [&]() {
            float v1 = -1.0f, v2 = -1.0f, v3 = -1.0f;
            u32 p1 = 0, p2 = 0, p3 = 0;
#pragma unroll 1
            for (u32 u = 0; u < NU; ++u) {
                const float4 v = pv[(warp * NU + u) * 32u + lane];
                const u32 l0 = wpos0 + (u * 32u + lane) * 4u;
                uint4 idv = make_uint4(0, 0, 0, 0);
                if constexpr (IDS) idv = pid[(warp * NU + u) * 32u + lane];
#pragma unroll
                for (int e4 = 0; e4 < 4; ++e4) {
                    const float x = f4get(v, e4);
                    const u32 px = IDS ? (e4 == 0 ? idv.x : e4 == 1 ? idv.y : e4 == 2 ? idv.z : idv.w) : l0 + e4;
                    // slots are visited in ascending position: strict '>' keeps the lowest position; virtual positions
                    // come in no order, so equal distances compare them
                    const bool g1 = x > v1 || (IDS && x == v1 && x >= 0.0f && px < p1);
                    const bool g2 = x > v2 || (IDS && x == v2 && x >= 0.0f && px < p2);
                    const bool g3 = x > v3 || (IDS && x == v3 && x >= 0.0f && px < p3);
                    v3 = g2 ? v2 : (g3 ? x : v3);
                    p3 = g2 ? p2 : (g3 ? px : p3);
                    v2 = g1 ? v1 : (g2 ? x : v2);
                    p2 = g1 ? p1 : (g2 ? px : p2);
                    v1 = g1 ? x : v1;
                    p1 = g1 ? px : p1;
                }
            }
            u64 k1 = v1 < 0.0f ? 0ull : make_key(v1, G_LOW - p1), k2 = v2 < 0.0f ? 0ull : make_key(v2, G_LOW - p2),
                k3 = v3 < 0.0f ? 0ull : make_key(v3, G_LOW - p3);
            u64 res[4] = {0ull, 0ull, 0ull, 0ull};
            u32 pops = 0;
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                const u64 wk = warp_max_key(k1);
                res[e] = wk;
                if ((FLAT || e < 2) && wk != 0ull && k1 == wk) {
                    k1 = k2;
                    k2 = k3;
                    k3 = 0ull;
                    ++pops;
                }
            }
            if constexpr (FLAT) {   // three candidates + the fourth key as the slice bound
                res[3] = warp_max_key(k1);
                // a lane popped three times no longer knows its next key: everything it holds sorts below its third,
                // so the third key itself is a valid (conservative) bound
                if (__any_sync(FULL, pops >= 3)) res[3] = res[2];
            }
            if (lane < 4) wtop[warp * 4 + lane] = lane == 0 ? res[0] : lane == 1 ? res[1] : lane == 2 ? res[2] : res[3];
            smax = __uint_as_float((u32)(res[0] >> 32));
        };

        // ---- the first pick: the point at POSITION start (wrapper.hpp:54-55) ------------------------------------------
        if (tid == 0) {
            // kd-line: the point at POSITION start; vanilla: the forced start indices, in order (lib.cpp:151-155)
            const u32 ns = IDS ? a.n_starts : 1u;
            for (u32 i = 0; i < ns; ++i) {
                u32 cur = a.starts ? (u32)a.starts[(size_t)cloud * ns + i] : 0u;
                if constexpr (IDS) cur = a.n - 1u - cur;
                tpos[i] = cur;
                tval[i] = FLT_MAX;
                for (u32 c = 0; c < DIM; ++c) tc[c * G_ECAP + i] = c < dim ? __ldg(qf + (size_t)c * npad + cur) : 0.0f;
                rel[i] = i;
                if (cta == 0) out[i] = (u64)cur;   // positions now, original ids by grid_map_kernel at the end
            }
            misc[M_NREL] = ns;
        }
        reselect();          // every valid point is at FLT_MAX: smax = FLT_MAX, so the first pick touches every slice
        __syncthreads();
        float clo[DIM], chi[DIM];
#pragma unroll
        for (int c = 0; c < DIM; ++c) {
            clo[c] = c < (int)dim ? ord2f(cbox[c]) : 0.0f;
            chi[c] = c < (int)dim ? ord2f(cbox[DIM + c]) : 0.0f;
        }

        u32 t = IDS ? a.n_starts : 1u;
#pragma unroll 1
        while (t < k) {
            const long long c0 = GCLK();
            // ================= A: apply the accepted picks that can touch this slice, re-select ======================
            {
                const u32 nrel = misc[M_NREL];
                bool dirty = false;
#pragma unroll 1
                for (u32 r0 = 0; r0 < nrel; r0 += 32) {
                    const u32 ri = (r0 + lane < nrel) ? rel[r0 + lane] : G_NONE;
                    bool touch = false;
                    if (ri != G_NONE) {
                        float pc[DIM];
#pragma unroll
                        for (int c = 0; c < DIM; ++c) pc[c] = tc[c * G_ECAP + ri];
                        touch = g_boxdist<DIM>(pc, wlo, whi) < smax;
                    }
                    const u32 mask = __ballot_sync(FULL, touch);
                    if (!mask) continue;
                    dirty = true;
#pragma unroll 1
                    for (u32 u = 0; u < NU; ++u) {
                        const u32 slot = (warp * NU + u) * 32u + lane;
                        float4 x[DIM];
#pragma unroll
                        for (int c = 0; c < DIM; ++c) x[c] = pts[(size_t)c * PCQ + slot];
                        float4 v = pv[slot];
                        u64 X01[DIM], X23[DIM];   // the four points of this slot, two per packed operand (FADD2 / FFMA2)
#pragma unroll
                        for (int c = 0; c < DIM; ++c) X01[c] = pk2(x[c].x, x[c].y), X23[c] = pk2(x[c].z, x[c].w);
                        u32 m = mask;
#pragma unroll 1
                        while (m) {
                            const u32 j = __ffs(m) - 1;
                            m &= m - 1;
                            const u32 rj = __shfl_sync(FULL, ri, j);
                            u64 RC[DIM];
#pragma unroll
                            for (int c = 0; c < DIM; ++c) {
                                const float r = tc[c * G_ECAP + rj];
                                RC[c] = pk2(r, r);
                            }
                            float d0, d1, d2, d3;
                            up2(sqdist2<DIM>(X01, RC, a.negzero), d0, d1);
                            up2(sqdist2<DIM>(X23, RC, a.negzero), d2, d3);
                            v.x = fminf(v.x, d0);   // std::min(dis, d), Point.h:82-86
                            v.y = fminf(v.y, d1);
                            v.z = fminf(v.z, d2);
                            v.w = fminf(v.w, d3);
                        }
                        pv[slot] = v;
                    }
                }
                if (dirty) {
                    __syncwarp();
                    reselect();
                }
            }
            __syncthreads();
            const long long c1 = GCLK();
            // ================= B: every warp ranks its own two keys among the CTA's 64 and publishes ranks 0..8 ==========
            const u32 stamp = (round + 1) & ((1u << (32 - G_PBITS)) - 1u);
            const u32 sbits = stamp << G_PBITS;
            float ctamax;
            if constexpr (FLAT) {   // every warp publishes its own four words: no CTA-level merge, no barrier
                ctamax = 0.0f;      // derived from the gathered keys below
                u64 *dst = reinterpret_cast<u64 *>(pubg) + (((size_t)(round & 1u) * G + cta) * G_W + warp) * WK;
                if (lane < WK) {
                    const u64 w = wtop[warp * 4 + lane];
                    stg_relaxed_v2(dst + lane, (u32)(w >> 32), (w ? G_LOW - (u32)w : G_PNONE) | sbits);
                }
            } else {
                const u64 a0 = wtop[lane * 4], a1 = wtop[lane * 4 + 1];
                ctamax = __uint_as_float(__reduce_max_sync(FULL, (u32)(a0 >> 32)));
                u64 *dst = reinterpret_cast<u64 *>(pubg) + ((size_t)(round & 1u) * G + cta) * G_NK;
#pragma unroll 1
                for (u32 s2 = 0; s2 < 2; ++s2) {
                    const u64 mk = wtop[warp * 4 + s2];
                    if (mk == 0ull) break;
                    const u32 r = __popc(__ballot_sync(FULL, a0 > mk)) + __popc(__ballot_sync(FULL, a1 > mk));
                    if (r > G_M) break;   // the second key sorts behind the first
                    if (lane == 0) stg_relaxed_v2(dst + r, (u32)(mk >> 32), (G_LOW - (u32)mk) | sbits);
                }
                if (warp == 0) {        // entries beyond the CTA's non-empty keys
                    const u32 nz = __popc(__ballot_sync(FULL, a0 != 0ull)) + __popc(__ballot_sync(FULL, a1 != 0ull));
                    if (lane >= nz && lane <= G_M) stg_relaxed_v2(dst + lane, 0u, G_PNONE | sbits);
                } else if (warp == 1) { // the largest slice bound
                    const u64 wb = warp_max_key(wtop[lane * 4 + 2]);
                    if (lane == 0) stg_relaxed_v2(dst + G_M + 1, (u32)(wb >> 32), (wb ? G_LOW - (u32)wb : G_PNONE) | sbits);
                }
            }
            const long long c2 = GCLK();
            // ================= C: gather every CTA's words (the grid-wide synchronisation): one 16-byte chunk per thread ====
            const u32 NCHK = FLAT ? G * G_W * (WK / 2) : G * (G_NK / 2);   // 16-byte chunks of one exchange buffer
            if (tid < NCHK) {
                const uint4 *src = pubg + (size_t)(round & 1u) * NCHK + tid;
                uint4 v;
                do {
                    v = ldg_relaxed_v4(src);
                } while ((v.y >> G_PBITS) != stamp || (v.w >> G_PBITS) != stamp);
                const u32 pa = v.y & G_PNONE, pb = v.w & G_PNONE;
                gk[2 * tid] = pa == G_PNONE ? 0ull : (((u64)v.x << 32) | (G_LOW - pa));
                gk[2 * tid + 1] = pb == G_PNONE ? 0ull : (((u64)v.z << 32) | (G_LOW - pb));
            }
            __syncwarp();   // lanes leave the spin loop one by one: converge before the warp collectives below
            if (tid < 16) wflag[tid] = 0;
            if (tid < 2) misc[M_E0 + tid] = 0;
            if (tid == 0) {
                misc[M_NREL] = 0;
            }
            __syncthreads();
            const long long c3 = GCLK();
            // ================= D: eligible candidates, sorted ============================================================
            // A candidate is eligible when it sorts above EVERY CTA's bound.  With the first M' keys of every CTA as
            // candidates, CTA c's bound is max(its slice bound, its key[M']); M' = 8, 4 or 1: the largest that leaves at
            // most ECAP candidates (three lists are built at once).
            const u32 NWG = (G + 31) >> 5;   // warps holding CTAs: thread c < G looks after CTA c's ten keys
            long long d1 = c3;
            if constexpr (FLAT) {
                // every warp of the group published {k0, k1, k2, bound}.  Candidates: the three keys of every warp against
                // Bd3 = the largest bound; if that leaves more than ECAP, only the first key of every warp against Bd1 =
                // the largest second key; if even that is too many (up to 512 warps), the single largest key -- always a
                // valid round.  Every warp derives the thresholds itself: no barrier.
                const u32 NWP = G * G_W;
                u64 b3 = 0ull, b1 = 0ull, b0 = 0ull;
                u32 cm = 0;
                for (u32 w = lane; w < NWP; w += 32) {
                    b3 = key_max(b3, gk[w * 4 + 3]);
                    b1 = key_max(b1, gk[w * 4 + 1]);
                    b0 = key_max(b0, gk[w * 4]);
                    if (w / G_W == cta) cm = max(cm, (u32)(gk[w * 4] >> 32));
                }
                const u64 Bd3 = warp_max_key(b3), Bd1 = warp_max_key(b1), Top = warp_max_key(b0);
                ctamax = __uint_as_float(__reduce_max_sync(FULL, cm));
#pragma unroll 1
                for (u32 idx0 = 0; idx0 < NWP * 3; idx0 += G_T) {
                    if (idx0 + warp * 32 >= NWP * 3) break;   // warp-uniform
                    const u32 idx = idx0 + tid, w = idx / 3, e = idx - w * 3;
                    const u64 kk = idx < NWP * 3 ? gk[w * 4 + e] : 0ull;
                    const bool f3 = kk > Bd3, f1 = e == 0 && kk > Bd1;
                    const u32 m3 = __ballot_sync(FULL, f3), m1 = __ballot_sync(FULL, f1);
                    u32 o3 = 0, o1 = 0;
                    if (lane == 0) {
                        if (m3) o3 = atomicAdd(&misc[M_E0], (u32)__popc(m3));
                        if (m1) o1 = atomicAdd(&misc[M_E0 + 1], (u32)__popc(m1));
                    }
                    const u32 lt = (1u << lane) - 1u;
                    o3 = __shfl_sync(FULL, o3, 0) + __popc(m3 & lt);
                    o1 = __shfl_sync(FULL, o1, 0) + __popc(m1 & lt);
                    if (f3 && o3 < G_ECAP) ekey[o3] = kk;
                    if (f1 && o1 < G_ECAP) ekey[G_ECAP + o1] = kk;
                }
                if (tid == 0) red[28] = Top;
            } else if (warp < NWG) {
                u64 kc[G_NK];
#pragma unroll
                for (u32 e = 0; e < G_NK; ++e) kc[e] = tid < G ? gk[tid * G_NK + e] : 0ull;
                const u64 wb = kc[G_M + 1];
                const u64 b8 = warp_max_key(key_max(wb, kc[8])), b4 = warp_max_key(key_max(wb, kc[4])),
                          b1 = warp_max_key(key_max(wb, kc[1]));
                if (lane == 0) {
                    red[warp * 4] = b8;
                    red[warp * 4 + 1] = b4;
                    red[warp * 4 + 2] = b1;
                }
                bar_sync_named(2, NWG * 32);
                d1 = GCLK();
                u64 Bd8 = 0ull, Bd4 = 0ull, Bd1 = 0ull;
#pragma unroll 1
                for (u32 w = 0; w < NWG; ++w) {
                    Bd8 = key_max(Bd8, red[w * 4]);
                    Bd4 = key_max(Bd4, red[w * 4 + 1]);
                    Bd1 = key_max(Bd1, red[w * 4 + 2]);
                }
                // the eligible keys of a CTA are a prefix of its descending list: three counts, packed into one word
                u32 c8 = 0, c4 = 0;
#pragma unroll
                for (u32 e = 0; e < G_M; ++e) {
                    c8 += kc[e] > Bd8;
                    if (e < 4) c4 += kc[e] > Bd4;
                }
                const u32 c1 = kc[0] > Bd1;
                const u32 mine = c8 | (c4 << 11) | (c1 << 22);   // every total is below 2^11
                u32 inc = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const u32 y = __shfl_up_sync(FULL, inc, o);
                    if (lane >= (u32)o) inc += y;
                }
                u32 *tot = reinterpret_cast<u32 *>(red + 24);   // [8] warp totals (red[0..19] hold the bounds)
                if (lane == 31) tot[warp] = inc;
                bar_sync_named(2, NWG * 32);
                u32 all = 0, before = 0;
#pragma unroll 1
                for (u32 w = 0; w < NWG; ++w) {
                    const u32 x = tot[w];
                    all += x;
                    before += w < warp ? x : 0u;
                }
                const u32 E8 = all & 0x7ffu, E4 = (all >> 11) & 0x7ffu;
                const u32 vs = E8 <= ECAP ? 0u : (E4 <= ECAP ? 1u : 2u);
                const u32 shf = vs * 11u;
                const u32 cnt = (mine >> shf) & 0x7ffu, off = ((before + inc - mine) >> shf) & 0x7ffu;
#pragma unroll
                for (u32 e = 0; e < G_M; ++e)
                    if (e < cnt && off + e < G_ECAP) ekey[off + e] = kc[e];
                if (tid == 0) misc[M_E0] = (all >> shf) & 0x7ffu;
            }
            __syncthreads();
            const long long d2 = GCLK();
            const u32 vflat = (FLAT && misc[M_E0] > ECAP) ? (misc[M_E0 + 1] > ECAP ? 2u : 1u) : 0u;
            const u32 E = vflat == 2 ? 1u : min(misc[M_E0 + vflat], G_ECAP);
            const u64 *ek = vflat == 2 ? red + 28 : ekey + vflat * G_ECAP;
            // rank by counting: 1024 / E2 adjacent lanes per candidate (E2 = E rounded up to a power of two >= 32); the
            // candidate's coordinates are fetched from the region (L2) meanwhile
            {
                const u32 sh = E <= 32 ? 5u : E <= 64 ? 4u : E <= 128 ? 3u : 2u;   // log2(lanes per candidate)
                const u32 np = 1u << sh, i = tid >> sh, part = tid & (np - 1u);
                const u32 gm = __ballot_sync(FULL, i < E);
                if (i < E) {   // whole groups of np lanes; groups never straddle a warp
                    const u64 myk = ek[i];
                    const u32 pos = G_LOW - (u32)myk;
                    float cc[DIM];
#pragma unroll
                    for (int c = 0; c < DIM; ++c) cc[c] = 0.0f;
                    if (part == 0) {
#pragma unroll
                        for (int c = 0; c < DIM; ++c)
                            if (c < (int)dim) cc[c] = __ldg(qf + (size_t)c * npad + pos);
                    }
                    u32 cn = 0;
#pragma unroll 2
                    for (u32 m = part; m < E; m += np) cn += ek[m] > myk;
#pragma unroll 1
                    for (u32 o = np >> 1; o; o >>= 1) cn += __shfl_xor_sync(gm, cn, o);
                    if (part == 0) {
                        tpos[cn] = pos;
                        tval[cn] = __uint_as_float((u32)(myk >> 32));
#pragma unroll
                        for (int c = 0; c < DIM; ++c) tc[c * G_ECAP + cn] = cc[c];
                    }
                }
            }
            __syncthreads();
            const long long c4 = GCLK();
            // ================= E: which candidates are the next picks ======================================================
            // In sorted order, candidate j is a pick unless an earlier PICK of this round lowers it (dist(P_j, P_i) < val_j);
            // a lowered candidate is skipped but raises the FLOOR every later pick must beat to its new key (its distance can
            // only fall further), next to every CTA's bound.  The first candidate at or below the floor ends the round.
            //   E1  conflict bits conf[j][i] for all i < j (warp per row, lanes split i, one ballot per 32 columns)
            //   E2  low[j] = exists i < j: conf[j][i] and not low[i]  -- depends on smaller indices only: iterated from the
            //       empty set it reaches its unique solution in (longest chain) steps, every step fully parallel
            //   E3  a lowered candidate's new key -> where the sorted list falls to it (binary search) -> stop = min of those
#pragma unroll 1
            for (u32 j = warp; j < E; j += G_W) {
                float pj[DIM];
#pragma unroll
                for (int c = 0; c < DIM; ++c) pj[c] = tc[c * G_ECAP + j];
                const float vj = tval[j];
                u32 any = 0;
#pragma unroll 1
                for (u32 i0 = 0; i0 < j; i0 += 32) {
                    const u32 i = i0 + lane;
                    float pi[DIM];
#pragma unroll
                    for (int c = 0; c < DIM; ++c) pi[c] = tc[c * G_ECAP + i];   // i < G_ECAP always
                    const u32 w = __ballot_sync(FULL, i < j && sqdist<DIM>(pj, pi) < vj);
                    if (lane == 0) conf[j * 8 + (i0 >> 5)] = w;
                    any |= w;
                }
                if (lane == 0) rowany[j] = any;
            }
            const long long g1 = GCLK();
            if (tid < 16) lw[tid] = 0;
            if (tid == 0) misc[M_STOP] = E;
            __syncthreads();
            long long g2 = g1, g3 = g1;
            u32 nit = 0;
            if (warp < 8) {   // E <= 256 candidates: one thread each, named barrier 1
                const u32 j = tid;
                const u32 nblk = (j + 31) >> 5;
                const bool has = j < E && rowany[j] != 0;
                u32 cur = 0;
#pragma unroll 1
                for (u32 it = 0;; ++it) {
                    bool low = false;
                    if (has) {
#pragma unroll 1
                        for (u32 b2 = 0; b2 < nblk; ++b2) low |= (conf[j * 8 + b2] & ~lw[cur * 8 + b2]) != 0;
                    }
                    const u32 bal = __ballot_sync(FULL, low);
                    if (lane == 0) {
                        lw[(cur ^ 1) * 8 + warp] = bal;
                        if (bal != lw[cur * 8 + warp]) wflag[it & 15] = 1;
                    }
                    if (tid == 0) wflag[(it + 1) & 15] = 0;
                    bar_sync_named(1, 256);
                    cur ^= 1;
                    ++nit;
                    if (!wflag[it & 15]) break;   // nothing changed: lw[cur] is the solution (uniform: written before the barrier)
                }
                // lw[cur] holds the lowered set
                g2 = GCLK();
                const bool low = j < E && ((lw[cur * 8 + warp] >> lane) & 1u);
                if (j < E) {
                    u32 st = E;
                    if (tval[j] == 0.0f) st = j;   // a zero-distance candidate ends the batch (see below)
                    if (low) {
                        float pj[DIM];
#pragma unroll
                        for (int c = 0; c < DIM; ++c) pj[c] = tc[c * G_ECAP + j];
                        float m = tval[j];
#pragma unroll 1
                        for (u32 b2 = 0; b2 < nblk; ++b2) {
                            u32 w = conf[j * 8 + b2] & ~lw[cur * 8 + b2];
                            while (w) {
                                const u32 i = b2 * 32 + (__ffs(w) - 1);
                                w &= w - 1;
                                float pi[DIM];
#pragma unroll
                                for (int c = 0; c < DIM; ++c) pi[c] = tc[c * G_ECAP + i];
                                m = fminf(m, sqdist<DIM>(pj, pi));
                            }
                        }
                        const u64 fk = make_key(m, G_LOW - tpos[j]);
                        u32 lo2 = j, hi2 = E;   // keys[0..j] > fk; find the first index whose key is <= fk
                        while (hi2 - lo2 > 1) {
                            const u32 mid = (lo2 + hi2) >> 1;
                            if (make_key(tval[mid], G_LOW - tpos[mid]) > fk) lo2 = mid;
                            else hi2 = mid;
                        }
                        st = min(st, hi2);
                    }
                    if (st < E) atomicMin(&misc[M_STOP], st);
                }
                g3 = GCLK();
                bar_sync_named(1, 256);   // M_STOP is final
                if (warp == 0) {          // picks = candidates before `stop` that are not lowered: mask words, their prefix counts
                    const u32 stop = misc[M_STOP];
                    u32 m = 0;
                    if (lane < 8) {
                        const u32 lim = stop > lane * 32 ? min(stop - lane * 32, 32u) : 0u;
                        m = ~lw[cur * 8 + lane] & (lim == 32 ? 0xffffffffu : ((1u << lim) - 1u));
                    }
                    u32 inc = __popc(m);
#pragma unroll
                    for (int o = 1; o < 8; o <<= 1) {
                        const u32 y = __shfl_up_sync(FULL, inc, o);
                        if (lane >= (u32)o) inc += y;
                    }
                    if (lane < 8) {
                        pickw[lane] = m;
                        pickw[8 + lane] = inc - __popc(m);
                    }
                    if (lane == 7) misc[M_J] = inc;
                }
            }
            __syncthreads();
            const long long c5 = GCLK();
            const bool allzero = tval[0] == 0.0f;   // every remaining distance is 0: the same position wins for ever
            u32 J = misc[M_J];
            if (allzero) J = 1;
            if (J > k - t) J = k - t;
            if (allzero) {
                if (cta == 0) {
                    const u64 p0 = (u64)tpos[0];
#pragma unroll 1
                    for (u32 x = t + tid; x < k; x += G_T) out[x] = p0;
                }
                t = k;
            } else {
                if (tid < G_ECAP && ((pickw[tid >> 5] >> (tid & 31)) & 1u)) {
                    const u32 idx = pickw[8 + (tid >> 5)] + __popc(pickw[tid >> 5] & ((1u << (tid & 31)) - 1u));
                    if (idx < J) {
                        if (cta == 0) out[t + idx] = (u64)tpos[tid];
                        // picks that can touch this CTA at all (CTA box against the CTA maximum)
                        float pj[DIM];
#pragma unroll
                        for (int c = 0; c < DIM; ++c) pj[c] = tc[c * G_ECAP + tid];
                        if (g_boxdist<DIM>(pj, clo, chi) < ctamax) rel[atomicAdd(&misc[M_NREL], 1u)] = tid;
                    }
                }
                t += J;
            }
            ++round;
            __syncthreads();
#if GDBG
            if (tid == 0 && round - 1 < 4096) g_grid_trace[(round - 1) * 160 + (blockIdx.x % 160)] = (u32)(c1 - c0);
            if (dbg_on) {
                const long long c6 = GCLK();
                dbg[0] += 1;
                dbg[1] += J;
                dbg[2] += (u64)(c1 - c0);   // apply + re-select (slowest warp of CTA 0)
                dbg[3] += (u64)(c2 - c1);   // rank + publish
                dbg[4] += (u64)(c3 - c2);   // gather (waits for the slowest CTA)
                dbg[5] += (u64)(c4 - c3);   // bounds, compaction, rank, table
                dbg[6] += (u64)(c5 - c4);   // pair checks
                dbg[7] += (u64)(c6 - c5);   // output + relevant list
                dbg[8] += (u64)(d1 - c3);   // D: bounds
                dbg[9] += (u64)(d2 - d1);   // D: compaction
                dbg[10] += (u64)(c4 - d2);  // D: rank + table
                dbg[12] += E;
                dbg[13] += (u64)(g1 - c4);   // E1 conflict matrix
                dbg[14] += (u64)(g2 - g1);   // E2 fixed point
                dbg[15] += (u64)(g3 - g2);   // E3 floors
                dbg[11] += nit;
            }
#endif
        }
        __syncthreads();
    }
#if GDBG
    if (dbg_on)
        for (int i = 0; i < 16; ++i) g_grid_dbg[i] = dbg[i];
#else
    if (tid == 0 && blockIdx.x == 0) {   // rounds of group 0 over its clouds, picks they resolved (bench.py: latency roofline)
        g_grid_dbg[0] = round;
        g_grid_dbg[1] = (u64)((a.B + ngrp - 1) / ngrp) * (a.k - 1);
    }
#endif
}

// positions -> original ids (src/wrapper.hpp:57-59), after the sampling kernel
__global__ void grid_map_kernel(u64 *out, const unsigned char *region, size_t region_stride, u32 B, u32 k, u32 dim, u32 npad,
                                u32 n_virtual) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * k) return;
    if (n_virtual) {   // vanilla: virtual position -> original index
        out[i] = (u64)(n_virtual - 1u - (u32)out[i]);
        return;
    }
    const u32 *perm = reinterpret_cast<const u32 *>(region + (i / k) * region_stride) + (size_t)(dim + 1) * npad;
    out[i] = (u64)__ldg(perm + (u32)out[i]);
}

// vanilla: qv[c][vp] = pts[n - 1 - vp][c]
__global__ void grid_reverse_kernel(const float *pts, float *qv, size_t qv_stride, u32 B, u32 n, u32 dim, u32 npad) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * n * dim) return;
    const u32 b = (u32)(i / ((size_t)n * dim));
    const size_t f = i - (size_t)b * n * dim;
    const u32 p = (u32)(f / dim), c = (u32)(f - (size_t)p * dim);
    qv[(size_t)b * qv_stride + (size_t)c * npad + (n - 1u - p)] = pts[i];
}

// ======================================================================================================
//  host side
// ======================================================================================================
static int pad_dim_g(int dim) { return dim <= 2 ? 2 : dim == 3 ? 3 : dim == 4 ? 4 : dim <= 6 ? 6 : 8; }

static size_t grid_smem(int dimp, u32 ppt, size_t S, bool flat, u32 gc = 16, bool ids = false) {
    const size_t pcq = (size_t)G_W * (ppt / 4) * 32;
    size_t b = (size_t)(dimp + 1) * pcq * 16;            // points
    b += flat ? (size_t)gc * G_W * 4 * 8 : (size_t)G_MAXG * G_NK * 8;  // gathered keys
    b += G_W * 4 * 8 + 2 * G_ECAP * 8 + 32 * 8;          // wtop, ekey, red
    b += (size_t)G_ECAP * 4 * 2;                         // tpos, tval
    b += (size_t)dimp * G_ECAP * 4 + G_ECAP * 4;         // tc, rel
    b += 2 * dimp * 4 + 16 * 4 + 16 * 4 + ((S + 1 + 3) & ~(size_t)3) * 4;   // cbox, misc, wflag, cum
    if (ids) b += pcq * 16 + 16;                          // virtual positions
    b += 32 * 4 + G_ECAP * 4 + G_ECAP * 8 * 4 + 1024 * 4;// lw, pickw, rowany, conf, slt
    return (b + 15) & ~(size_t)15;
}

size_t kd_grid_pub_bytes(const GridPlan &pl) {
    return pl.flat ? (size_t)pl.groups * 2 * pl.gc * G_W * 4 * 8 : (size_t)pl.groups * 2 * G_MAXG * G_NK * 8;
}

bool plan_kdline_grid(size_t n, size_t dim, size_t h, size_t B, int n_sms, GridPlan *pl, bool ids) {
    if (dim == 0 || dim > 8 || h == 0 || n == 0 || B == 0 || n >= G_PNONE) return false;
    const int want = tuning().grid;
    if (want == 0) return false;
    const int dimp = pad_dim_g((int)dim);
    const size_t sms = n_sms < (int)G_MAXG ? n_sms : G_MAXG;
    const size_t S = (size_t)1 << (h < 20 ? h : 20);
    pl->dimp = dimp;
    pl->ecap = G_ECAP;
    if (const int v = tuning().grid_ecap; v >= 32 && v <= (int)G_ECAP) pl->ecap = (u32)v;
    // ---- a batch of medium clouds: groups of gc <= 8 CTAs per cloud, every warp publishes its own keys (flat mode) ------
    const int grp = tuning().group;
    const bool medium = n >= 8192 && n < 262144 && S <= 4096;
    // measured (scripts/cmp_group.py): 1.4x - 3x faster than the cluster coordinator/worker kernel and than the one-warp-per-
    // cloud kernel on every shape from 8192 points up (FPS_B200_GROUP=0 switches it off)
    const bool group_default = true;
    if (grp != 0 && (grp == 1 || (want < 0 && medium && group_default)) && want != 1) {
        for (u32 p = 12; p >= 4; p -= 4) {
            const size_t per_cta = (size_t)G_T * p;
            for (u32 gc = 1; gc <= 16; gc *= 2) {
                if (grid_smem(dimp, p, S, true, gc, ids) > 227 * 1024) break;
                // room for the subtree packing: ~20 % slack (a tighter fit falls back to position ranges in the kernel)
                if (gc * per_cta * 5 < n * 6) continue;
                if (gc > sms) break;
                size_t groups = sms / gc;
                if (groups > B) groups = B;
                // worth it when the groups keep most of the GPU busy or the clouds are few
                pl->ppt = p;
                pl->gc = gc;
                pl->groups = (u32)groups;
                pl->G = (u32)(groups * gc);
                pl->flat = 1;
                pl->smem = grid_smem(dimp, p, S, true, gc, ids);
                return true;
            }
        }
        if (grp == 1) return false;
    }
    // slices never straddle a leaf: at most ceil(n / SL) + S of them (every leaf ends with one partial slice)
    u32 ppt = 0, G = 0;
    for (u32 p = 4; p <= 16 && !ppt; p += 4) {
        const size_t SL = 32 * (size_t)p, slices = (n + SL - 1) / SL + S;
        const size_t g = (slices + G_W - 1) / G_W;
        if (g <= sms && grid_smem(dimp, p, S, false, 16, ids) <= 227 * 1024) {
            ppt = p;
            G = (u32)g;
        }
    }
    if (!ppt) return false;
    // a few huge clouds: as many at once as whole groups of G CTAs fit; measured against the cluster coordinator/worker
    // kernel (scripts/cmp_mid.py) it wins from ~130 k points as long as the batch needs at most two passes
    size_t groups = sms / G;
    if (groups > B) groups = B;
    if (groups < 1) groups = 1;
    if (want < 0 && (n < 131072 || (B + groups - 1) / groups > 2)) return false;
    pl->ppt = ppt;
    pl->G = (u32)(groups * G);
    pl->gc = G;
    pl->groups = (u32)groups;
    pl->flat = 0;
    pl->smem = grid_smem(dimp, ppt, S, false, 16, ids);
    return true;
}

template <int DIM, bool FLAT, bool IDS>
static cudaError_t launch_grid_t(const GridPlan &pl, GridArgs &a, cudaStream_t st) {
    auto kern = kdline_grid_kernel<DIM, FLAT, IDS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) return e;
    void *params[] = {&a};
    // cooperative launch: every CTA must be resident, the rounds synchronise through global memory
    return cudaLaunchCooperativeKernel((void *)kern, dim3(pl.G), dim3(G_T), params, pl.smem, st);
}

cudaError_t grid_debug_counters(u64 *out16) {
#if GDBG
    static u32 tr[4096 * 160];
    if (cudaMemcpyFromSymbol(tr, g_grid_trace, sizeof(tr)) == cudaSuccess) {
        u64 rounds = 0;
        cudaMemcpyFromSymbol(&rounds, g_grid_dbg, 8);
        double smax = 0, smean = 0;
        u32 worst = 0;
        const u32 R = rounds < 4096 ? (u32)rounds : 4096u;
        for (u32 r = 0; r < R; ++r) {
            u32 mx = 0;
            double sm = 0;
            for (u32 c = 0; c < 144; ++c) {
                mx = tr[r * 160 + c] > mx ? tr[r * 160 + c] : mx;
                sm += tr[r * 160 + c];
            }
            smax += mx;
            smean += sm / 144;
            worst = mx > worst ? mx : worst;
        }
        if (R) fprintf(stderr, "  grid trace: phase A per round: mean over CTAs %.0f cycles, max over CTAs %.0f (worst round %u)\n", smean / R, smax / R, worst);
    }
#endif
    return cudaMemcpyFromSymbol(out16, g_grid_dbg, sizeof(u64) * 16);
}

template <bool FLAT, bool IDS>
static cudaError_t launch_grid_d(const GridPlan &pl, GridArgs &a, cudaStream_t st) {
    switch (pl.dimp) {
        case 2: return launch_grid_t<2, FLAT, IDS>(pl, a, st);
        case 3: return launch_grid_t<3, FLAT, IDS>(pl, a, st);
        case 4: return launch_grid_t<4, FLAT, IDS>(pl, a, st);
        case 6: return launch_grid_t<6, FLAT, IDS>(pl, a, st);
        default: return launch_grid_t<8, FLAT, IDS>(pl, a, st);
    }
}

cudaError_t launch_kdline_grid(const GridPlan &pl, unsigned char *region, size_t region_stride, const u64 *starts, u64 *out,
                               unsigned char *pub, u32 B, u32 n, u32 dim, u32 k, u32 h, cudaStream_t st, const float *pts_vanilla,
                               float *qv, u32 n_starts) {
    GridArgs a;
    a.negzero = 0x8000000080000000ull;
    a.S = 1u << h;
    a.region = region;
    a.region_stride = region_stride;
    a.starts = starts;
    a.out = out;
    a.pub = reinterpret_cast<uint4 *>(pub);
    a.B = B;
    a.n = n;
    a.npad = (n + 31) & ~31u;
    a.dim = dim;
    a.k = k;
    a.ppt = pl.ppt;
    a.ecap = pl.ecap;
    a.gc = pl.gc;
    const bool ids = pts_vanilla != nullptr;
    a.qv = qv;
    a.qv_stride = (size_t)dim * a.npad;
    a.n_starts = n_starts ? n_starts : 1;
    cudaError_t e = cudaMemsetAsync(pub, 0, kd_grid_pub_bytes(pl), st);
    if (e != cudaSuccess) return e;
    if (ids) {
        const size_t tot = (size_t)B * n * dim;
        grid_reverse_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(pts_vanilla, qv, a.qv_stride, B, n, dim, a.npad);
        count_launch();
    }
    if (pl.flat) e = ids ? launch_grid_d<true, true>(pl, a, st) : launch_grid_d<true, false>(pl, a, st);
    else e = ids ? launch_grid_d<false, true>(pl, a, st) : launch_grid_d<false, false>(pl, a, st);
    count_launch();
    if (e != cudaSuccess) return e;
    const size_t tot = (size_t)B * k;
    grid_map_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(out, region, region_stride, B, k, dim, a.npad, ids ? n : 0u);
    count_launch();
    return cudaGetLastError();
}

}  // namespace fps
