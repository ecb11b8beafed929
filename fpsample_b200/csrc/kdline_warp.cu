// kdline_warp.cu -- QuickFPS kd-line SAMPLING for clouds that fit on chip: ONE WARP PER CLOUD, no block barrier
// and no cross-warp traffic on the pick loop.  Many clouds are resident per SM: a cloud's points live either in
// shared memory or in TENSOR MEMORY (TMEM, 256 KB per SM) used as a lane-private, dynamically indexed
// scratchpad through tcgen05.ld / tcgen05.st.  Position p belongs to lane p % 32 and chunk p / 32; both stores
// are "lane-major": a lane's values of one component for consecutive chunks are contiguous, so 8 chunks (256
// positions) of one component move with one tcgen05.ld.x8 / two 128-bit shared loads per lane.  At 4096 points x
// 3 dims a cloud is 64 KB: 3 clouds in shared memory + 4 in TMEM = 7 warps = 7 clouds per SM, 1036 per B200, so
// BASELINE.json's cfg 2 (1024 clouds) runs in a single wave.
//
// Semantics (reference src/_ext/KDLineTree.h:56-85, src/_ext/KDNode.h:84-166, src/wrapper.hpp:54-59):
// exact FPS over the array the kd build permuted (built by kdline.cu / kdbuild.cu into the per-cloud region),
// started at POSITION start, running distance initialised to FLT_MAX (src/_ext/Point.h:61-65), ties to the
// lowest position (strict '>' everywhere).  Buckets (kd leaves) follow the reference's own lazy scheme
// (KDNode::update_distance, KDNode.h:120-166): lane b owns bucket b -- box, current max, the max point's
// coordinates -- and per new sample either drops it (box bound >= max, KDNode.h:105-118), defers it to the
// bucket's pending list (max point not affected, KDNode.h:124-134) or flushes: one pass over the bucket applies
// every pending sample and recomputes the max (KDNode.h:147-161).  A full pending list flushes early, which is
// always allowed.  Float rounding is monotone, so skipped work never changes a distance and the result equals the
// eager recurrence bit for bit (tests/test_oracle.py::test_lazy_equals_eager, GPU parity suite).
#include <cfloat>
#include <type_traits>

#include "common.cuh"
#include "engine.h"

namespace fps {

constexpr u32 W_NONE = 0xffffffffu;
constexpr int W_U = 8;                 // chunks (of 32 positions) per straight-line block of a bucket pass
constexpr u32 W_MAX_SMEM_WARPS = 12;
constexpr u32 W_TMEM_COLS = 512;
constexpr u32 W_MAXR = 12;             // pending samples per bucket at most

#ifndef WDBG
#define WDBG 0   // 1: per-phase clock64 counters of cloud 0 (scripts/run_one.py prints them)
#endif
__device__ u64 g_warp_dbg[16];

struct WarpArgs {
    unsigned char *region;
    size_t region_stride;
    const u64 *starts;
    u64 *out;
    u32 *counter;
    u32 B, n, npad, dim, k, S, nlo_pad;
    u32 nch;            // chunks per cloud, padded to a multiple of W_U
    u32 n_tmem_warps, n_smem_warps, slot_bytes, meta_bytes, R, lazy, hybrid;
};

__device__ __forceinline__ float4 lds128(u32 a) {
    float4 f;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f.x), "=f"(f.y), "=f"(f.z), "=f"(f.w) : "r"(a));
    return f;
}
__device__ __forceinline__ void sts128(u32 a, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

__device__ __forceinline__ float2 lds64(u32 a) {
    float2 f;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(f.x), "=f"(f.y) : "r"(a));
    return f;
}
__device__ __forceinline__ void sts64(u32 a, float x, float y) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(x), "f"(y) : "memory");
}

// ---- the two point stores ---------------------------------------------------------------------------------------
// component c (0..DIM-1 coordinates, DIM = running distance) of position (chunk, lane)
struct SmemStore {
    static constexpr bool kTrackCoords = false;
    u32 base;   // shared-space byte address of this warp's slot
    u32 lst;    // words per (component, lane) row: nch + 4 (16-byte aligned rows, conflict-free 128-bit access)
    __device__ __forceinline__ u32 addr(u32 comp, u32 lane, u32 chunk) const { return base + ((comp * 32u + lane) * lst + chunk) * 4u; }
    // U consecutive chunks: 8 start at a multiple of 4 (two 128-bit accesses), 4 and 6 at an even chunk (64-bit accesses)
    template <int U>
    __device__ __forceinline__ void load(u32 comp, u32 lane, u32 cb, float (&v)[U]) const {
        const u32 a = addr(comp, lane, cb);
        if constexpr (U == 8) {
            const float4 f0 = lds128(a), f1 = lds128(a + 16u);
            v[0] = f0.x, v[1] = f0.y, v[2] = f0.z, v[3] = f0.w, v[4] = f1.x, v[5] = f1.y, v[6] = f1.z, v[7] = f1.w;
        } else {
#pragma unroll
            for (int i = 0; i < U; i += 2) {
                const float2 f = lds64(a + 4u * i);
                v[i] = f.x, v[i + 1] = f.y;
            }
        }
    }
    __device__ __forceinline__ void wait_ld() const {}
    template <int U>
    __device__ __forceinline__ void store(u32 comp, u32 lane, u32 cb, const float (&v)[U]) const {
        const u32 a = addr(comp, lane, cb);
        if constexpr (U == 8) {
            sts128(a, v[0], v[1], v[2], v[3]);
            sts128(a + 16u, v[4], v[5], v[6], v[7]);
        } else {
#pragma unroll
            for (int i = 0; i < U; i += 2) sts64(a + 4u * i, v[i], v[i + 1]);
        }
    }
    __device__ __forceinline__ void wait_st() const {}
    // coordinates of position p, to every lane
    template <int DIM>
    __device__ __forceinline__ void load_point(u32 p, u32, float (&r)[DIM]) const {
#pragma unroll
        for (int c = 0; c < DIM; ++c) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r[c]) : "r"(addr(c, p & 31u, p >> 5)));
    }
};

struct TmemStore {
    // a point lookup is three tcgen05.ld + a wait + three shuffles on the pick path: lanes remember their candidate's
    // coordinates instead (three selects per chunk of a bucket pass) and the winner broadcasts them
    static constexpr bool kTrackCoords = true;
    u32 base;   // tensor-memory address of this warp's lane quarter: (32 * (warp % 4)) << 16 | first column
    u32 nch;    // columns per component
    __device__ __forceinline__ void ld4(u32 col, u32 *w) const {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(col) : "memory");
    }
    __device__ __forceinline__ void st4(u32 col, const float *v) const {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(col), "r"(__float_as_uint(v[0])),
                     "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3]))
                     : "memory");
    }
    template <int U>
    __device__ __forceinline__ void load(u32 comp, u32, u32 cb, float (&v)[U]) const {
        u32 w[U];
        const u32 col = base + comp * nch + cb;
        if constexpr (U == 8) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                         : "r"(col)
                         : "memory");
        } else {
            ld4(col, w);
            if constexpr (U == 6)
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(w[4]), "=r"(w[5]) : "r"(col + 4u) : "memory");
        }
#pragma unroll
        for (int i = 0; i < U; ++i) v[i] = __uint_as_float(w[i]);
    }
    __device__ __forceinline__ void wait_ld() const { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
    template <int U>
    __device__ __forceinline__ void store(u32 comp, u32, u32 cb, const float (&v)[U]) const {
        const u32 col = base + comp * nch + cb;
        if constexpr (U == 8) {
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(col),
                         "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                         "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                         : "memory");
        } else {
            st4(col, v);
            if constexpr (U == 6)
                asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(col + 4u), "r"(__float_as_uint(v[4])),
                             "r"(__float_as_uint(v[5]))
                             : "memory");
        }
    }
    __device__ __forceinline__ void wait_st() const { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
    // the point lives in lane p % 32 of this warp's quarter: every lane reads its own column, the owner broadcasts
    template <int DIM>
    __device__ __forceinline__ void load_point(u32 p, u32, float (&r)[DIM]) const {
        u32 w[DIM];
#pragma unroll
        for (int c = 0; c < DIM; ++c)
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(w[c]) : "r"(base + c * nch + (p >> 5)) : "memory");
        wait_ld();
#pragma unroll
        for (int c = 0; c < DIM; ++c) r[c] = __uint_as_float(__shfl_sync(FULL, w[c], p & 31u));
    }
};

// Coordinates in shared memory, running distances in tensor memory: a cloud of up to 16 384 points x 3 dims
// (BASELINE.json cfg 3) is 192 KB of coordinates + 64 KB of distances -- more than either store alone, exactly what one
// SM has when both are used.  One such warp per SM (its distances fill one lane quarter of TMEM).
struct HybridStore {
    static constexpr bool kTrackCoords = false;
    SmemStore s;   // DIM components
    TmemStore t;   // one component: the distance of chunk c is column c
    u32 dimc;      // index of the distance component (= DIM of the kernel)
    template <int U>
    __device__ __forceinline__ void load(u32 comp, u32 lane, u32 cb, float (&v)[U]) const {
        if (comp == dimc) t.load<U>(0, lane, cb, v);
        else s.load<U>(comp, lane, cb, v);
    }
    __device__ __forceinline__ void wait_ld() const { t.wait_ld(); }
    template <int U>
    __device__ __forceinline__ void store(u32 comp, u32 lane, u32 cb, const float (&v)[U]) const {
        if (comp == dimc) t.store<U>(0, lane, cb, v);
        else s.store<U>(comp, lane, cb, v);
    }
    __device__ __forceinline__ void wait_st() const { t.wait_st(); }
    template <int DIM>
    __device__ __forceinline__ void load_point(u32 p, u32 lane, float (&r)[DIM]) const {
        s.load_point(p, lane, r);
    }
};

// point -> box squared distance (KDNode.h:105-118) without branches: the excess along a dimension is
// max(r - hi, lo - r, 0) -- the same subtraction result the reference's if/else picks, or (+-)0
template <int DIM>
__device__ __forceinline__ float boxdist_nb(const float (&r)[DIM], const float (&lo)[DIM], const float (&hi)[DIM]) {
    float acc = 0.0f;
#pragma unroll
    for (int j = 0; j < DIM; ++j) {
        const float e = fmaxf(fmaxf(__fsub_rn(r[j], hi[j]), __fsub_rn(lo[j], r[j])), 0.0f);
        const float e2 = __fmul_rn(e, e);
        acc = (j == 0) ? e2 : __fadd_rn(acc, e2);
    }
    return acc;
}

// ---- one cloud on one warp --------------------------------------------------------------------------------------
template <int DIM, int BPL, class ST>
__device__ __forceinline__ void warp_cloud(const WarpArgs &a, const ST st, u32 cloud, u32 *nlo_s, u32 pend /* shared addr */) {
    constexpr u32 PRB = ((DIM + 3) / 4) * 16;   // bytes per pending-list entry
    const u32 lane = lane_id();
    const u32 n = a.n, npad = a.npad, dim = a.dim, S = a.S, k = a.k, nch = a.nch, R = a.R;
    unsigned char *rg = a.region + (size_t)cloud * a.region_stride;
    const float *q = reinterpret_cast<const float *>(rg);
    const u32 *perm = reinterpret_cast<const u32 *>(rg) + (size_t)(dim + 1) * npad;
    const u32 *nlo = perm + npad;
    const float *fbox = reinterpret_cast<const float *>(nlo + a.nlo_pad);
    u64 *out = a.out + (size_t)cloud * k;

    // ---- stage the permuted cloud into this warp's store; distances start at FLT_MAX (Point.h:61-65) ---------
    for (u32 cb = 0; cb < nch; cb += W_U) {
#pragma unroll
        for (int c = 0; c <= DIM; ++c) {
            float v[W_U];
#pragma unroll
            for (int u = 0; u < W_U; ++u) {
                const u32 p = (cb + u) * 32 + lane;
                v[u] = (c == DIM) ? FLT_MAX : ((c < (int)dim && p < n) ? __ldg(q + (size_t)c * npad + p) : 0.0f);
            }
            st.template store<W_U>(c, lane, cb, v);
        }
    }
    for (u32 s = lane; s <= S; s += 32) nlo_s[s] = nlo[s];
    st.wait_st();
    __syncwarp();

    // ---- bucket state in registers: lane l owns buckets l, l+32, ... ---------------------------------------------
    float blo[BPL][DIM], bhi[BPL][DIM], bmc[BPL][DIM], bmax[BPL];
    u32 bpos[BPL], np[BPL];
    u32 valid = 0;
#pragma unroll
    for (int j = 0; j < BPL; ++j) {
        const u32 b = j * 32 + lane;
        bmax[j] = FLT_MAX;   // every bucket flushes on the first sample (KDNode::init, KDNode.h:84-103)
        bpos[j] = 0;
        np[j] = 0;
#pragma unroll
        for (int c = 0; c < DIM; ++c) blo[j][c] = bhi[j][c] = bmc[j][c] = 0.0f;
        if (b < S && nlo_s[b + 1] > nlo_s[b]) {
            valid |= 1u << j;
#pragma unroll
            for (int c = 0; c < DIM; ++c)
                if (c < (int)dim) {
                    blo[j][c] = fbox[(size_t)b * 2 * dim + c];
                    bhi[j][c] = fbox[(size_t)b * 2 * dim + dim + c];
                }
        }
    }

    u32 cur = a.starts ? (u32)a.starts[cloud] : 0u;   // POSITION in the permuted array (wrapper.hpp:54-55)
    float r[DIM];
#pragma unroll
    for (int c = 0; c < DIM; ++c) r[c] = (c < (int)dim) ? __ldg(q + (size_t)c * npad + cur) : 0.0f;
    u32 mypos = cur;   // lane (t % 32) remembers pick t until the block of 32 picks is written out
    u32 idv = 0;       // ids of the block being written out

#if WDBG
    u64 dbg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
    for (u32 t = 1; t < k; ++t) {
#if WDBG
        const long long c0 = clock64();
#endif
        // ---- every bucket against the new sample: drop / defer / flush (KDNode.h:120-146) -------------------------
        u32 mask[BPL];
#pragma unroll
        for (int j = 0; j < BPL; ++j) {
            const bool ok = (valid >> j) & 1u;
            const bool touch = boxdist_nb<DIM>(r, blo[j], bhi[j]) < bmax[j];     // can lower something in the bucket
            const bool hitmax = !(sqdist<DIM>(bmc[j], r) > bmax[j]);             // lowers the bucket's max point
            const bool want = ok && (touch || hitmax);
            if (want) {   // remember the sample: pend[np][bucket]
                const u32 e = pend + (np[j] * S + (u32)(j * 32) + lane) * PRB;
                sts128(e, r[0], DIM > 1 ? r[DIM > 1 ? 1 : 0] : 0.f, DIM > 2 ? r[DIM > 2 ? 2 : 0] : 0.f, DIM > 3 ? r[DIM > 3 ? 3 : 0] : 0.f);
                if constexpr (DIM > 4)
                    sts128(e + 16u, r[4], DIM > 5 ? r[DIM > 5 ? 5 : 0] : 0.f, DIM > 6 ? r[DIM > 6 ? 6 : 0] : 0.f, 0.f);
                ++np[j];
            }
            const bool flush = want && (a.lazy ? (hitmax || np[j] >= R) : true);
            mask[j] = __ballot_sync(FULL, flush);
        }
        __syncwarp();   // pending entries are read by every lane below
#if WDBG
        const long long c1 = clock64();
#pragma unroll
        for (int j = 0; j < BPL; ++j) dbg[5] += __popc(mask[j]);
#endif
        // ---- flush: one pass over the bucket applies all its pending samples (KDNode.h:147-161) -----------------------
        for (;;) {
            u32 b = W_NONE;
#pragma unroll
            for (int j = 0; j < BPL; ++j) {
                if (b == W_NONE && mask[j]) {
                    b = j * 32 + (__ffs(mask[j]) - 1);
                    mask[j] &= mask[j] - 1;
                }
            }
            if (b == W_NONE) break;
            u32 npj = 0;
#pragma unroll
            for (int j = 0; j < BPL; ++j)
                if ((b >> 5) == (u32)j) npj = np[j];
            const u32 nref = __shfl_sync(FULL, npj, b & 31u);
            const u32 lo = nlo_s[b], hi = nlo_s[b + 1], span = hi - lo;
            const u32 c1b = (hi - 1) >> 5;
            float best = -1.0f;
            u32 bi = W_NONE;
            float bc[DIM];   // coordinates of this lane's best point (global store only)
#pragma unroll
            for (int c = 0; c < DIM; ++c) bc[c] = 0.0f;
            // one straight-line block of U chunks (U x 32 positions): all pending samples applied, first maximum per lane
            auto block = [&](auto Uc, const u32 cb) {
                constexpr int U = decltype(Uc)::value;
                float x[DIM][U], v[U], old[U];
#pragma unroll
                for (int c = 0; c < DIM; ++c) st.template load<U>(c, lane, cb, x[c]);
                st.template load<U>(DIM, lane, cb, old);
                st.wait_ld();
#pragma unroll
                for (int u = 0; u < U; ++u) v[u] = old[u];
#if WDBG
                dbg[6] += nref;
#endif
                // the next pending sample is fetched while the current one is applied
                const u32 e0 = pend + b * PRB, estep = S * PRB;
                float4 n0 = lds128(e0), n1 = make_float4(0.f, 0.f, 0.f, 0.f);
                if constexpr (DIM > 4) n1 = lds128(e0 + 16u);
                // (kept rolled: the compiler's 2-way unrolling with its entry paths costs 1.7 % -- measured, 1.225 -> 1.204 ms on cfg 2)
#pragma unroll 1
                for (u32 i = 0; i < nref; ++i) {
                    const float4 f0 = n0;
                    [[maybe_unused]] const float4 f1 = n1;
                    const u32 en = e0 + min(i + 1, nref - 1) * estep;
                    n0 = lds128(en);
                    if constexpr (DIM > 4) n1 = lds128(en + 16u);
                    float w[8];
                    w[0] = f0.x, w[1] = f0.y, w[2] = f0.z, w[3] = f0.w;
                    if constexpr (DIM > 4) w[4] = f1.x, w[5] = f1.y, w[6] = f1.z, w[7] = f1.w;
                    float ref[DIM];
#pragma unroll
                    for (int c = 0; c < DIM; ++c) ref[c] = w[c];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        float pt[DIM];
#pragma unroll
                        for (int c = 0; c < DIM; ++c) pt[c] = x[c][u];
                        v[u] = fminf(v[u], sqdist<DIM>(pt, ref));   // std::min(dis, d), Point.h:82-86
                    }
                }
                // positions outside [lo, hi) belong to a neighbour bucket (or are padding): they keep their value
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const u32 p = (cb + u) * 32 + lane;
                    const bool in = (p - lo) < span;
                    v[u] = in ? v[u] : old[u];
                    if (in && v[u] > best) {   // ascending p: a lane keeps its first maximum
                        best = v[u];
                        bi = p;
                        if constexpr (ST::kTrackCoords) {
#pragma unroll
                            for (int c = 0; c < DIM; ++c) bc[c] = x[c][u];
                        }
                    }
                }
                st.template store<U>(DIM, lane, cb, v);
            };
            // the bucket's chunks from an even chunk on (64-bit aligned rows): most buckets of ~128 points fit 6 chunks,
            // small ones 4; nch is a multiple of 8, so a block pulled back from the end stays aligned
            const u32 c0 = (lo >> 5) & ~1u, ncn = c1b - c0 + 1;
            if (ncn <= 4) {
                block(std::integral_constant<int, 4>{}, min(c0, nch - 4u));
            } else if (ncn <= 6) {
                block(std::integral_constant<int, 6>{}, min(c0, nch - 6u));
            } else {
                u32 cb0 = c0 & ~3u;
                for (; cb0 <= c1b; cb0 += W_U) block(std::integral_constant<int, W_U>{}, min(cb0, nch - W_U));
            }
            // a neighbouring bucket may share this bucket's first / last chunk.  (Waiting here, right behind the stores, is
            // faster than waiting in front of the next pass's loads: 1.23 against 1.31 ms on BASELINE cfg 2.)
            st.wait_st();
            // bucket max, then its lowest position; the owner lane takes both plus the point's coordinates
            const float pv = fmaxf(best, 0.0f);
            const u32 m = __reduce_max_sync(FULL, __float_as_uint(pv));
            const u32 qpos = __reduce_min_sync(FULL, (__float_as_uint(pv) == m) ? bi : W_NONE);
            float mc[DIM];
            if constexpr (ST::kTrackCoords) {
#pragma unroll
                for (int c = 0; c < DIM; ++c) mc[c] = __shfl_sync(FULL, bc[c], qpos & 31u);
            } else {
                st.load_point(qpos, lane, mc);
            }
            if (lane == (b & 31u)) {
#pragma unroll
                for (int j = 0; j < BPL; ++j)
                    if ((b >> 5) == (u32)j) {
                        bmax[j] = __uint_as_float(m);
                        bpos[j] = qpos;
                        np[j] = 0;
#pragma unroll
                        for (int c = 0; c < DIM; ++c) bmc[j][c] = mc[c];
                    }
            }
        }
#if WDBG
        const long long c2 = clock64();
#endif
        // ---- arg-max over buckets: largest max, lowest position (KDLineTree.h:56-67) -------------------------------
        u32 kmax = 0, cand = W_NONE;
#pragma unroll
        for (int j = 0; j < BPL; ++j) {
            if ((valid >> j) & 1u) {
                const u32 kb = __float_as_uint(bmax[j]);
                if (cand == W_NONE || kb > kmax || (kb == kmax && bpos[j] < cand)) {
                    kmax = kb;
                    cand = bpos[j];
                }
            }
        }
        const u32 M = __reduce_max_sync(FULL, kmax);
        const u32 mine = (cand != W_NONE && kmax == M) ? cand : W_NONE;
        cur = __reduce_min_sync(FULL, mine);
        if constexpr (ST::kTrackCoords) {   // the owner lane holds the max point's coordinates
            float cm[DIM];
#pragma unroll
            for (int c = 0; c < DIM; ++c) cm[c] = 0.0f;
#pragma unroll
            for (int j = 0; j < BPL; ++j)
                if (((valid >> j) & 1u) && bpos[j] == cur) {
#pragma unroll
                    for (int c = 0; c < DIM; ++c) cm[c] = bmc[j][c];
                }
            const u32 src = __ffs(__ballot_sync(FULL, mine == cur)) - 1;   // positions are unique: exactly one lane
#pragma unroll
            for (int c = 0; c < DIM; ++c) r[c] = __shfl_sync(FULL, cm[c], src);
        } else {
            st.load_point(cur, lane, r);
        }
        // ---- output: positions are turned into original ids 32 picks at a time (wrapper.hpp:57-59) -----------------
        // the id lookup of a finished block is issued one pick before its result is stored: its L2 latency hides
        // behind that pick's work
        if ((t & 31u) == 0) idv = __ldg(perm + mypos);
        if ((t & 31u) == 1 && t > 1) out[t - 33 + lane] = (u64)idv;
        if (lane == (t & 31u)) mypos = cur;
#if WDBG
        const long long c3 = clock64();
        dbg[0] += 1;
        dbg[1] += (u64)(c1 - c0);
        dbg[2] += (u64)(c2 - c1);
        dbg[4] += (u64)(c3 - c2);
        dbg[7] += (u64)(c3 - c0);
#endif
    }
#if WDBG
    if (cloud == 0 && lane == 0)
        for (int i = 0; i < 8; ++i) g_warp_dbg[i] = dbg[i];
#endif
    if (k > 32 && ((k - 1) & 31u) == 0) out[k - 33 + lane] = (u64)idv;   // the block whose lookup the last pick issued
    {   // tail: picks [k0, k)
        const u32 k0 = (k - 1) & ~31u;
        if (k0 + lane < k) out[k0 + lane] = (u64)__ldg(perm + mypos);
    }
    __syncwarp();
}

template <int DIM, int BPL>
__global__ void __launch_bounds__(512, 1) kdline_warp_kernel(WarpArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ u32 tbase_s;
    const u32 warp = warp_id(), lane = lane_id();
    const u32 nw = a.hybrid ? a.n_tmem_warps : a.n_tmem_warps + a.n_smem_warps;

    if (a.n_tmem_warps) {   // whole tensor memory of this SM: one CTA per SM by construction (see plan)
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase_s)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    // per-warp metadata: pending lists [R][S] entries | nlo_s[S + 1] (bucket boundaries)
    unsigned char *meta = smem_raw + (size_t)warp * a.meta_bytes;
    const u32 pend = smem_u32(meta);
    u32 *nlo_s = reinterpret_cast<u32 *>(meta + (size_t)a.R * a.S * ((DIM + 3) / 4) * 16);
    unsigned char *slots = smem_raw + (size_t)nw * a.meta_bytes;

    // first wave: cloud = warp * gridDim + cta (few clouds spread over SMs before they stack up on one);
    // afterwards clouds are handed out dynamically
    u32 cloud = warp * gridDim.x + blockIdx.x;
    while (cloud < a.B) {
        if (a.hybrid) {   // warp w: coordinates in shared-memory slot w, distances in TMEM lane quarter w
            HybridStore st;
            st.s.base = smem_u32(slots + (size_t)warp * a.slot_bytes);
            st.s.lst = a.nch + 4;
            st.t.base = tbase_s + ((warp * 32u) << 16);
            st.t.nch = a.nch;
            st.dimc = DIM;
            warp_cloud<DIM, BPL>(a, st, cloud, nlo_s, pend);
        } else if (warp < a.n_tmem_warps) {
            TmemStore st;
            st.base = tbase_s + ((warp * 32u) << 16);
            st.nch = a.nch;
            warp_cloud<DIM, BPL>(a, st, cloud, nlo_s, pend);
        } else {
            SmemStore st;
            st.base = smem_u32(slots + (size_t)(warp - a.n_tmem_warps) * a.slot_bytes);
            st.lst = a.nch + 4;
            warp_cloud<DIM, BPL>(a, st, cloud, nlo_s, pend);
        }
        u32 nxt = 0;
        if (lane == 0) nxt = atomicAdd(a.counter, 1u);
        cloud = __shfl_sync(FULL, nxt, 0) + nw * gridDim.x;
    }
    if (a.n_tmem_warps) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase_s) : "memory");
    }
}

// ======================================================================================================
//  host side
// ======================================================================================================
static int warp_dim(int dim) { return dim <= 2 ? 2 : dim == 3 ? 3 : dim == 4 ? 4 : dim <= 6 ? 6 : 7; }

bool plan_kdline_warp(size_t n, size_t dim, size_t h, size_t B, int n_sms, WarpPlan *pl) {
    if (dim == 0 || dim > 7 || h == 0 || h > 7 || n == 0 || B == 0) return false;
    const Tuning &tu = tuning();
    if (tu.warp == 0) return false;
    const size_t S = (size_t)1 << h;
    const int dimp = warp_dim((int)dim);
    const size_t nch = (((n + 31) / 32) + W_U - 1) / W_U * W_U;
    const size_t slot = (size_t)(dimp + 1) * 32 * (nch + 4) * 4;
    u32 tm = ((size_t)(dimp + 1) * nch <= W_TMEM_COLS) ? 4u : 0u;
    if (tu.warp_tmem == 0) tm = 0;
    const bool lazy = tu.warp_lazy != 0;
    const size_t cap = 227 * 1024 - 64;
    const size_t pr = (size_t)((dimp + 3) / 4) * 16;
    auto meta_of = [&](size_t R) { return (R * S * pr + (S + 1) * 4 + 15) & ~(size_t)15; };
    // as many shared-memory clouds as fit with the smallest useful pending lists, then the lists grow into the rest
    const size_t Rmin = lazy ? 3 : 1;
    u32 sw = 0;
    while (sw < W_MAX_SMEM_WARPS && (sw + 1 + tm) * meta_of(Rmin) + (sw + 1) * slot <= cap) ++sw;
    if (S > 32 && dimp > 4) sw = tm = 0;   // on chip, 4 buckets per lane only with small records (register budget)
    pl->hybrid = 0;
    if (sw + tm == 0 && !(S > 32 && dimp > 4) && nch <= W_TMEM_COLS) {
        // neither store alone holds the cloud: coordinates in shared memory + distances in TMEM (one lane quarter each)
        const size_t cslot = (size_t)dimp * 32 * (nch + 4) * 4;
        u32 hw = 0;
        while (hw < 4 && (hw + 1) * (meta_of(Rmin) + cslot) <= cap) ++hw;
        // measured on cfg 3 (64 x 16384 x 3 -> 4096, h=7): 6.7 ms against 5.8 ms for the async cluster kernel (4 buckets
        // per lane + TMEM round trips make the lone warp's pick ~2900 cycles), so it is opt-in: FPS_B200_WARP_HYBRID=1
        const bool hyb = hw > 0 && tu.warp_hybrid == 1;
        if (hyb) {
            size_t R = Rmin;
            while (lazy && R < W_MAXR && hw * (meta_of(R + 1) + cslot) <= cap) ++R;
            size_t grid = B < (size_t)n_sms ? B : (size_t)n_sms;
            pl->dimp = dimp;
            pl->rs = (u32)R;
            pl->bpl = S <= 32 ? 1 : 4;
            pl->n_tmem_warps = hw;
            pl->n_smem_warps = 0;
            pl->slot_bytes = (u32)cslot;
            pl->meta_bytes = (u32)meta_of(R);
            pl->grid = (u32)grid;
            pl->lazy = lazy ? 1 : 0;
            pl->nch = (u32)nch;
            pl->hybrid = 1;
            pl->smem = hw * (meta_of(R) + cslot);
            if (pl->smem < 120 * 1024) pl->smem = 120 * 1024;
            return true;
        }
    }
    if (sw + tm == 0) return false;   // not on chip: the streaming sampler (kdline_stream.cu) or the cluster / grid kernels take it
    const size_t nw = sw + tm;
    size_t R = Rmin;
    while (lazy && R < W_MAXR && nw * meta_of(R + 1) + sw * slot <= cap) ++R;
    size_t grid = (B + nw - 1) / nw;
    if (grid < (size_t)n_sms) grid = B < (size_t)n_sms ? B : (size_t)n_sms;   // spread before stacking
    if (grid > (size_t)n_sms) grid = n_sms;
    pl->dimp = dimp;
    pl->rs = (u32)R;
    pl->bpl = S <= 32 ? 1 : 4;
    pl->n_tmem_warps = tm;
    pl->n_smem_warps = sw;
    pl->slot_bytes = (u32)slot;
    pl->meta_bytes = (u32)meta_of(R);
    pl->grid = (u32)grid;
    pl->lazy = lazy ? 1 : 0;
    pl->nch = (u32)nch;
    pl->smem = nw * meta_of(R) + sw * slot;
    // tensor memory is allocated whole: never let a second CTA of this kernel become resident on the SM
    if (tm && pl->smem < 120 * 1024) pl->smem = 120 * 1024;
    return true;
}

template <int DIM, int BPL>
static cudaError_t launch_warp_t(const WarpPlan &pl, const WarpArgs &a, cudaStream_t st) {
    auto kern = kdline_warp_kernel<DIM, BPL>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) return e;
    kern<<<pl.grid, 32 * (pl.n_tmem_warps + pl.n_smem_warps), pl.smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t warp_debug_counters(u64 *out16) { return cudaMemcpyFromSymbol(out16, g_warp_dbg, sizeof(u64) * 16); }

cudaError_t launch_kdline_warp(const WarpPlan &pl, unsigned char *region, size_t region_stride, const u64 *starts, u64 *out,
                               u32 *counter, u32 B, u32 n, u32 dim, u32 k, u32 h, cudaStream_t st) {
    WarpArgs a;
    a.region = region;
    a.region_stride = region_stride;
    a.starts = starts;
    a.out = out;
    a.counter = counter;
    a.B = B;
    a.n = n;
    a.npad = (n + 31) & ~31u;
    a.dim = dim;
    a.k = k;
    a.S = 1u << h;
    a.nlo_pad = (a.S + 1 + 31) & ~31u;
    a.nch = pl.nch;
    a.n_tmem_warps = pl.n_tmem_warps;
    a.n_smem_warps = pl.n_smem_warps;
    a.slot_bytes = pl.slot_bytes;
    a.meta_bytes = pl.meta_bytes;
    a.R = pl.rs;
    a.lazy = pl.lazy;
    a.hybrid = pl.hybrid;
    cudaError_t e = cudaMemsetAsync(counter, 0, 256, st);
    if (e != cudaSuccess) return e;
    const bool b1 = pl.bpl == 1;
    switch (pl.dimp) {
        case 2: e = b1 ? launch_warp_t<2, 1>(pl, a, st) : launch_warp_t<2, 4>(pl, a, st); break;
        case 3: e = b1 ? launch_warp_t<3, 1>(pl, a, st) : launch_warp_t<3, 4>(pl, a, st); break;
        case 4: e = b1 ? launch_warp_t<4, 1>(pl, a, st) : launch_warp_t<4, 4>(pl, a, st); break;
        case 6: e = launch_warp_t<6, 1>(pl, a, st); break;
        default: e = launch_warp_t<7, 1>(pl, a, st); break;
    }
    count_launch();
    return e;
}

}  // namespace fps
