// kdline_stream.cu -- QuickFPS kd-line SAMPLING for big clouds in big batches (BASELINE.json cfg 5: thousands of
// 100 000-point clouds): the points stay in the per-cloud region in global memory (HBM / L2) and a TEAM of 1, 2 or 4
// warps samples one cloud.  The governing roofline is HBM bandwidth: a bucket pass reads D coordinates + the running
// distance of every point of the bucket and writes back the distances that changed -- 4(D+2) bytes per point at most.
//
// Semantics (reference src/_ext/KDLineTree.h:56-85, src/_ext/KDNode.h:84-166, src/wrapper.hpp:54-59):
// exact FPS over the array the kd build permuted, started at POSITION start, running distance initialised to FLT_MAX
// (src/_ext/Point.h:61-65), ties to the lowest position (strict '>' everywhere).  Buckets (kd leaves) follow the
// reference's own lazy scheme (KDNode::update_distance, KDNode.h:120-166): a lane owns a bucket -- box, current max, the
// max point's coordinates -- and per new sample either drops it (box bound >= max, KDNode.h:105-118), defers it to the
// bucket's pending list (max point not affected, KDNode.h:124-134) or flushes: one pass over the bucket applies every
// pending sample and recomputes the max (KDNode.h:147-161).  A full pending list flushes early, which is always
// allowed (the sample is a genuine earlier pick).  Float rounding is monotone, so skipped work never changes a
// distance and the result equals the eager recurrence bit for bit.
//
// What is different from one warp per cloud (kdline_warp.cu, on-chip clouds):
//   * a bucket pass is split over the team's warps (each takes a contiguous run of 32-position chunks), so a pass is ONE
//     HBM round trip instead of three or four dependent ones -- what a small shard of a multi-GPU batch needs
//     (3.5 clouds per SM at 512 clouds: latency-bound);
//   * loads and stores are predicated per position: nothing outside [lo, hi) of the bucket is read, only distances that
//     changed are written -- DRAM traffic is the algorithmic traffic (ncu: 1.35x before);
//   * a lane owns four consecutive positions of a 128-position chunk: 128-bit loads and stores, a quarter of the memory
//     instructions and address arithmetic of a 32-bit-per-lane layout;
//   * executed-work counters (points scanned, point-updates, flushes, tests) for the roofline (SURVEY.md 8(d) W_exec).
// Per pick the team meets at two named barriers (bar.sync id, 32*WPC): after the tests (pending entries + the list of
// buckets to pass over) and after the bucket passes (per-warp partial maxima); merging those and the arg-max over all
// buckets are done by every warp for itself.  A bucket that is going to be passed over is prefetched into L2 with one
// cp.async.bulk.prefetch per component the moment its owner lane decides so, before the first barrier.
//
// Every warp of a team executes the whole pick loop, and a warp issues a dependent instruction every 4-5 cycles: the
// cost of a pick is its instruction count.  What keeps that count down ("Why the streaming fraction ..."):
//   * one base address per block and predicated 128-bit loads / stores (no divergent branch around memory accesses);
//   * exact-dimension kernels (EX: the cloud has exactly DIM dimensions; no per-component checks, no counters);
//   * the thread id read once through an opaque instruction (no S2R re-reads, no re-derived lane / team ids);
//   * flushed buckets claim list slots with a shared-memory counter (no ballots, popc prefix counts or searches);
//   * maxima travel as 64-bit keys (max bits, ~position): one compare per record, an empty slot is the smallest key;
//   * cross-lane results through redux.sync (22 cycles) rather than ballot + find-first-set + shuffle (~90).
// The host side cuts a batch into full waves of two-warp teams and a last, partial wave of four-warp teams
// (plan_kdline_stream; fps_b200_describe_stream_plan shows the cut).
#include <cfloat>

#include <cstdio>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "engine.h"

namespace fps {

constexpr u32 S_NONE = 0xffffffffu;
constexpr u32 S_MAXF = 4;           // flushed buckets per exchange batch (1.3 - 2.2 per pick on the BASELINE clouds)
#ifndef S_THREADS_DEF
#define S_THREADS_DEF 512
#endif
constexpr u32 S_THREADS = S_THREADS_DEF;   // 16 warps per CTA, one CTA per SM (128 registers per thread)
constexpr u32 S_REC = 48;           // bytes of a partial record: max bits, position, up to 8 coordinates
constexpr u32 S_MAXR = 16;          // pending samples per bucket at most

__device__ u64 g_stream_wexec[16];

struct StreamArgs {
    unsigned char *region;
    size_t region_stride;
    const u64 *starts;
    u64 *out;
    u32 *counter;       // dynamic cloud scheduler
    u32 B, n, npad, dim, k, S, nlo_pad, R, team_bytes, count, prefetch;
    u64 negzero;        // two binary32 -0.0 as an operand the compiler cannot see through (packed products, common.cuh)
};

template <int WPC>
__device__ __forceinline__ void team_sync(u32 team) {
    if constexpr (WPC == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(team + 1u), "n"(WPC * 32) : "memory");
}

__device__ __forceinline__ float4 s_lds128(u32 a) {
    float4 f;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f.x), "=f"(f.y), "=f"(f.z), "=f"(f.w) : "r"(a));
    return f;
}
__device__ __forceinline__ void s_sts128(u32 a, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

// point -> box squared distance (KDNode.h:105-118) without branches: the excess along a dimension is
// max(r - hi, lo - r, 0) -- the same subtraction result the reference's if/else picks, or (+-)0
template <int DIM>
__device__ __forceinline__ float s_boxdist(const float (&r)[DIM], const float (&lo)[DIM], const float (&hi)[DIM]) {
    float acc = 0.0f;
#pragma unroll
    for (int j = 0; j < DIM; ++j) {
        const float e = fmaxf(fmaxf(__fsub_rn(r[j], hi[j]), __fsub_rn(lo[j], r[j])), 0.0f);
        const float e2 = __fmul_rn(e, e);
        acc = (j == 0) ? e2 : __fadd_rn(acc, e2);
    }
    return acc;
}

__device__ __forceinline__ u32 lds32(u32 a) {
    u32 v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 lds128u(u32 a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts32(u32 a, u32 v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts128u(u32 a, u32 x, u32 y, u32 z, u32 w) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
// 128-bit loads under a per-lane predicate, as ONE predicated instruction each (an `if` around a load is a divergent branch
// with its reconvergence bookkeeping and a second copy of the fill values): lanes whose predicate is off keep `fill`
__device__ __forceinline__ float4 ldg128_if(const float *p, bool on, float fill) {   // read-only path (coordinates)
    float4 v = make_float4(fill, fill, fill, fill);
    asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %5, 0;\n @p ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n}"
                 : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
                 : "l"(p), "r"((u32)on));
    return v;
}
__device__ __forceinline__ void stcg128_if(float *p, const float4 &v, bool on) {
    asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %5, 0;\n @p st.global.cg.v4.f32 [%4], {%0, %1, %2, %3};\n}" ::"f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w), "l"(p), "r"((u32)on)
                 : "memory");
}
__device__ __forceinline__ void stcg32_if(float *p, float v, bool on) {
    asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %2, 0;\n @p st.global.cg.f32 [%1], %0;\n}" ::"f"(v), "l"(p), "r"((u32)on) : "memory");
}
__device__ __forceinline__ float4 ldcg128_if(const float *p, bool on, float fill) {   // L2 only (distances: rewritten all the time)
    float4 v = make_float4(fill, fill, fill, fill);
    asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %5, 0;\n @p ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];\n}"
                 : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
                 : "l"(p), "r"((u32)on)
                 : "memory");
    return v;
}
__device__ __forceinline__ float2 s_lds64(u32 a) {
    float2 f;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(f.x), "=f"(f.y) : "r"(a));
    return f;
}
__device__ __forceinline__ void s_sts64(u32 a, float x, float y) { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(x), "f"(y) : "memory"); }

// a pending-list entry = the sample's DIM coordinates: 16 bytes up to 4 dimensions, 24 for 6 (three 64-bit accesses: a third
// more entries per bucket in the same shared memory, and 6-D clouds defer ~14 samples per pick), 32 for 8
template <int DIM>
struct PendEntry {
    static constexpr u32 kBytes = DIM <= 4 ? 16u : DIM <= 6 ? 24u : 32u;
    static __device__ __forceinline__ void store(u32 e, const float (&r)[DIM]) {
        if constexpr (DIM <= 4) {
            s_sts128(e, r[0], DIM > 1 ? r[DIM > 1 ? 1 : 0] : 0.f, DIM > 2 ? r[DIM > 2 ? 2 : 0] : 0.f, DIM > 3 ? r[DIM > 3 ? 3 : 0] : 0.f);
        } else if constexpr (DIM <= 6) {
            s_sts64(e, r[0], r[1]);
            s_sts64(e + 8u, r[2], r[3]);
            s_sts64(e + 16u, r[4], DIM > 5 ? r[DIM > 5 ? 5 : 0] : 0.f);
        } else {
            s_sts128(e, r[0], r[1], r[2], r[3]);
            s_sts128(e + 16u, r[4], r[5], DIM > 6 ? r[DIM > 6 ? 6 : 0] : 0.f, DIM > 7 ? r[DIM > 7 ? 7 : 0] : 0.f);
        }
    }
    static __device__ __forceinline__ void load(u32 e, float4 &f0, float4 &f1) {
        f1 = make_float4(0.f, 0.f, 0.f, 0.f);
        if constexpr (DIM <= 4) {
            f0 = s_lds128(e);
        } else if constexpr (DIM <= 6) {
            const float2 a = s_lds64(e), b = s_lds64(e + 8u), c = s_lds64(e + 16u);
            f0 = make_float4(a.x, a.y, b.x, b.y);
            f1 = make_float4(c.x, c.y, 0.f, 0.f);
        } else {
            f0 = s_lds128(e);
            f1 = s_lds128(e + 16u);
        }
    }
};

__device__ __forceinline__ void l2_prefetch_bulk(const void *p, u32 bytes) {   // 16-byte aligned address, size a multiple of 16
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// One cloud on a team of WPC warps.  A lane owns FOUR consecutive positions of every 128-position chunk (128-bit loads and
// stores: a warp access is 512 contiguous bytes per component); a bucket pass is cut into runs of chunks, one run per warp.
// Per pick the team meets at TWO named barriers: after the tests (pending entries + the list of buckets to pass over) and
// after the passes (one partial maximum per warp and bucket).  Everything after that is done by every warp for itself --
// merging the partial maxima into its own copy of the bucket maxima, the arg-max over them -- so no third exchange is needed
// (checked with compute-sanitizer racecheck, profiles/r02_sanitizer_racecheck.log).
template <int DIM, int WPC, int BPL, bool EX>
__device__ __forceinline__ void stream_cloud(const StreamArgs &a, u32 cloud, u32 team, u32 tw, u32 lane, u32 tm /* shared address */, u32 cnt_s) {
    constexpr u32 SP = 32u * WPC * BPL;            // bucket slots of the team (>= S)
    constexpr u32 PRB = ((DIM + 3) / 4) * 16;      // bytes per max-point record
    constexpr u32 PE = PendEntry<DIM>::kBytes;     // bytes per pending-list entry
    constexpr u32 NW = WPC * BPL;                  // 32-bucket groups: flush-mask words, table slots per lane
    constexpr int G = 2;                           // 128-position chunks per block: 8 positions per lane in registers
    const u32 npad = a.npad, dim = a.dim, S = a.S, k = a.k, R = a.R;
    unsigned char *rg = a.region + (size_t)cloud * a.region_stride;
    const float *q = reinterpret_cast<const float *>(rg);
    float *dis = reinterpret_cast<float *>(rg) + (size_t)dim * npad;

    // team shared memory (32-bit shared addresses).  Everything whose size is known at compile time comes FIRST, so that its
    // address is `tm` + an immediate (no address arithmetic, no registers); the pending lists, whose depth R is a run-time
    // choice of the planner, come last:
    // bucket records [SP] {lo, hi, pending, -} | max-point coordinates [SP] | fcnt[2] (+pad) | flist[SP] (flushed buckets, compacted
    // per group) | part[2][S_MAXF][WPC] | fslot[WPC] | pending lists [R][SP]
    const u32 brec = tm;
    const u32 bmcs = brec + SP * 16;
    const u32 fcnt = bmcs + SP * PRB;   // two counters (pick parity)
    const u32 flist = fcnt + ((NW + 3) & ~3u) * 4;
    const u32 part = flist + SP * 4;
    const u32 fslot = part + 2 * S_MAXF * WPC * S_REC + tw * PRB;   // one per warp
    const u32 pend = part + 2 * S_MAXF * WPC * S_REC + WPC * PRB;

    // ---- distances start at FLT_MAX (Point.h:61-65); bucket boundaries to shared memory ------------------------------
    for (u32 p = (tw * 32 + lane) * 4; p < npad; p += WPC * 128)
        __stcg(reinterpret_cast<float4 *>(dis + p), make_float4(FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX));
    {
        const u32 *nlo = reinterpret_cast<const u32 *>(rg) + (size_t)(dim + 2) * npad;
        for (u32 s = tw * 32 + lane; s < SP; s += WPC * 32) {
            const u32 lo = nlo[s < S ? s : S], hi = nlo[s + 1 < S ? s + 1 : S];
            sts128u(brec + s * 16, lo, hi, 0u, 0u);
        }
    }

    // ---- the buckets this lane OWNS (tests, pending list): warp tw, slot j, lane l own bucket (tw * BPL + j) * 32 + l ------
    float blo[BPL][DIM], bhi[BPL][DIM], bmc[BPL][DIM];
    float bmax_[WPC == 1 ? 1 : BPL];   // (a one-warp team owns every bucket: its table below IS the owned state)
    u32 np[BPL];
    // ---- every warp's own copy of ALL bucket maxima: lane l, slot s = bucket s * 32 + l ----------------------------------------
    float tmax[NW];
    u32 tposn[NW], tvalid = 0;   // tposn: the maximum's position, COMPLEMENTED: (max bits, ~position) orders as one 64-bit key, and
                                 // an empty slot (0, 0) is the smallest key there is
#define S_OWNMAX(j) (*(WPC == 1 ? &tmax[(j)] : &bmax_[WPC == 1 ? 0 : (j)]))
    if (tw == 0 && lane == 0) sts32(fcnt, 0u), sts32(fcnt + 4, 0u);
    team_sync<WPC>(team);
    {
        const float *fbox = reinterpret_cast<const float *>(rg) + (size_t)(dim + 2) * npad + a.nlo_pad;
#pragma unroll
        for (int j = 0; j < BPL; ++j) {
            const u32 b = (tw * BPL + j) * 32 + lane;
            np[j] = 0;
#pragma unroll
            for (int c = 0; c < DIM; ++c) blo[j][c] = bhi[j][c] = bmc[j][c] = 0.0f;
            const uint4 br = lds128u(brec + b * 16);
            if (b < S && br.y > br.x) {
#pragma unroll
                for (int c = 0; c < DIM; ++c)
                    if (c < (int)dim) {
                        blo[j][c] = fbox[(size_t)b * 2 * dim + c];
                        bhi[j][c] = fbox[(size_t)b * 2 * dim + dim + c];
                    }
            }
        }
#pragma unroll
        for (u32 s = 0; s < NW; ++s) {
            const u32 b = s * 32 + lane;
            const uint4 br = lds128u(brec + b * 16);
            tmax[s] = 0.0f;
            tposn[s] = 0u;
            if (b < S && br.y > br.x) tvalid |= 1u << s;
        }
#pragma unroll
        for (int j = 0; j < BPL; ++j)   // every bucket flushes on the first sample (KDNode::init, KDNode.h:84-103); empty slots stay 0
            S_OWNMAX(j) = ((tvalid >> (tw * BPL + j)) & 1u) ? FLT_MAX : 0.0f;
    }

    u32 cur = a.starts ? (u32)a.starts[cloud] : 0u;   // POSITION in the permuted array (wrapper.hpp:54-55)
    float r[DIM];
#pragma unroll
    for (int c = 0; c < DIM; ++c) r[c] = (c < (int)dim) ? __ldg(q + (size_t)c * npad + cur) : 0.0f;
    u32 mypos = cur;   // lane (t % 32) remembers pick t until the block of 32 picks is written out
    u32 idv = 0;       // ids of the block being written out
    u32 bp = 0;        // parity of the partial-record buffer

    for (u32 t = 1; t < k; ++t) {
        // ---- 1. every owned bucket against the new sample: drop / defer / flush (KDNode.h:120-146) ---------------------------
#pragma unroll
        for (int j = 0; j < BPL; ++j) {
            const u32 grp = tw * BPL + j, b = grp * 32 + lane;
            const bool ok = (tvalid >> grp) & 1u;
            const bool touch = s_boxdist<DIM>(r, blo[j], bhi[j]) < S_OWNMAX(j);   // can lower something in the bucket
            const bool hitmax = !(sqdist<DIM>(bmc[j], r) > S_OWNMAX(j));          // lowers the bucket's max point
            const bool want = ok && (touch || hitmax);
            if (want) {   // remember the sample: pend[np][bucket]
                PendEntry<DIM>::store(pend + (np[j] * SP + b) * PE, r);
                ++np[j];
                sts32(brec + b * 16 + 8, np[j]);
            }
            const bool flush = want && (hitmax || np[j] >= R);
            if (flush) {   // into the team's list (any order will do: a pick's passes are independent of each other); the bucket's
                           // lines start moving from HBM to L2 right away
                const u32 slot = atomicAdd(reinterpret_cast<u32 *>(__cvta_shared_to_generic(fcnt + (t & 1u) * 4)), 1u);
                sts32(flist + slot * 4, b);
                if (a.prefetch) {
                    const uint4 br = lds128u(brec + b * 16);
                    const u32 p0 = br.x & ~3u, bytes = (((br.y + 3u) & ~3u) - p0) * 4u;
                    l2_prefetch_bulk(dis + p0, bytes);
#pragma unroll
                    for (int c = 0; c < DIM; ++c)
                        if (c < (int)dim) l2_prefetch_bulk(q + (size_t)c * npad + p0, bytes);
                }
            }
            if (!EX && a.count) {
                const u32 ne = __popc(__ballot_sync(FULL, flush && !hitmax));
                if (lane == 0 && ne) atomicAdd(reinterpret_cast<u64 *>(__cvta_shared_to_generic(cnt_s + 24)), (u64)ne);
            }
        }
        team_sync<WPC>(team);

        // ---- 2. bucket passes: every flushed bucket, this warp's run of chunks, all pending samples applied -----------------
        u64 fk = 0;                // the best maximum among the buckets passed over in THIS pick (as a record key) and its position; its
        u32 fq = S_NONE;           // point is parked in this WARP's own slot: if it wins the arg-max, the coordinates come from there
        const u32 total = lds32(fcnt + (t & 1u) * 4);                  // flushed buckets of this pick
        if (tw == 0 && lane == 0) sts32(fcnt + ((t + 1u) & 1u) * 4, 0u);   // the other counter: next pick's tests come after the next barrier
        for (u32 f0 = 0; f0 < total; f0 += S_MAXF) {   // batches of at most S_MAXF buckets between two exchanges
            const u32 nf = min(total - f0, S_MAXF);
            u32 myb = S_NONE;                         // lane fi remembers the bucket of flush slot fi
            const u32 pbuf = part + bp * (S_MAXF * WPC * S_REC);
            for (u32 fi = 0; fi < nf; ++fi) {
                const u32 f = f0 + fi;
                const u32 b = lds32(flist + f * 4);
                if (lane == fi) myb = b;
                const uint4 br = lds128u(brec + b * 16);
                const u32 lo = br.x, hi = br.y, nref = br.z;
                const u32 c0 = lo >> 7, ncn = ((hi - 1) >> 7) - c0 + 1, per = (ncn + WPC - 1) / WPC;
                const u32 my0 = c0 + tw * per, my1 = min(my0 + per, c0 + ncn);
                if (!EX && a.count && tw == 0 && lane == 0) {
                    u64 *cs = reinterpret_cast<u64 *>(__cvta_shared_to_generic(cnt_s));
                    atomicAdd(cs + 0, (u64)(hi - lo));
                    atomicAdd(cs + 1, (u64)(hi - lo) * nref);
                    atomicAdd(cs + 2, 1ull);
                }
                float best = -1.0f;
                u32 bi = S_NONE;
                float bc[DIM];
#pragma unroll
                for (int c = 0; c < DIM; ++c) bc[c] = 0.0f;
                for (u32 cb = my0; cb < my1; cb += G) {   // one block: G chunks of 128 positions, 4 per lane each
                    float4 x[DIM][G], old[G];
                    bool inside[G], edge = false;
                    // one base address per block: component c of group g lies c * npad + g * 128 floats behind it
                    const float *base = q + (size_t)(cb * 128 + lane * 4);
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const u32 p4 = (cb + g) * 128 + lane * 4;
                        const bool any = cb + g < my1 && p4 < hi && p4 + 3 >= lo;
                        inside[g] = cb + g < my1 && p4 >= lo && p4 + 3 < hi;
                        edge = edge || (any && !inside[g]);
                        old[g] = ldcg128_if(base + (size_t)dim * npad + g * 128, any, -1.0f);   // -1: never a maximum, never stored
#pragma unroll
                        for (int c = 0; c < DIM; ++c) x[c][g] = ldg128_if(base + (size_t)c * npad + g * 128, any && (EX || c < (int)dim), 0.0f);
                    }
                    const bool hasedge = __any_sync(FULL, edge);
                    if (hasedge) {   // a run's first / last group: positions of the neighbour buckets drop out
#pragma unroll
                        for (int g = 0; g < G; ++g) {
                            const u32 p4 = (cb + g) * 128 + lane * 4;
                            if (p4 + 0 < lo || p4 + 0 >= hi) old[g].x = -1.0f;
                            if (p4 + 1 < lo || p4 + 1 >= hi) old[g].y = -1.0f;
                            if (p4 + 2 < lo || p4 + 2 >= hi) old[g].z = -1.0f;
                            if (p4 + 3 < lo || p4 + 3 >= hi) old[g].w = -1.0f;
                        }
                    }
                    float4 v[G];
#pragma unroll
                    for (int g = 0; g < G; ++g) v[g] = old[g];
                    // the next pending sample is fetched while the current one is applied
                    const u32 e0 = pend + b * PE, estep = SP * PE;
                    float4 n0, n1;
                    PendEntry<DIM>::load(e0, n0, n1);
                    // (not unrolled: the compiler's 4-way unrolling with its entry paths costs 3-5 % -- measured)
#pragma unroll 1
                    for (u32 i = 0; i < nref; ++i) {
                        const float4 f0v = n0, f1v = n1;
                        PendEntry<DIM>::load(e0 + min(i + 1, nref - 1) * estep, n0, n1);
                        const float w8[8] = {f0v.x, f0v.y, f0v.z, f0v.w, f1v.x, f1v.y, f1v.z, f1v.w};
                        u64 RC[DIM];   // the sample in both halves of a packed operand (FADD2 / FFMA2: two positions per instruction)
#pragma unroll
                        for (int c = 0; c < DIM; ++c) RC[c] = pk2(w8[c], w8[c]);
#pragma unroll
                        for (int g = 0; g < G; ++g) {
                            u64 PA[DIM], PB[DIM];
#pragma unroll
                            for (int c = 0; c < DIM; ++c) PA[c] = pk2(x[c][g].x, x[c][g].y), PB[c] = pk2(x[c][g].z, x[c][g].w);
                            float d0, d1, d2, d3;
                            up2(sqdist2<DIM>(PA, RC, a.negzero), d0, d1);
                            up2(sqdist2<DIM>(PB, RC, a.negzero), d2, d3);
                            v[g].x = fminf(v[g].x, d0);   // std::min(dis, d), Point.h:82-86
                            v[g].y = fminf(v[g].y, d1);
                            v[g].z = fminf(v[g].z, d2);
                            v[g].w = fminf(v[g].w, d3);
                        }
                    }
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const u32 p4 = (cb + g) * 128 + lane * 4;
                        const bool ch = v[g].x != old[g].x || v[g].y != old[g].y || v[g].z != old[g].z || v[g].w != old[g].w;
                        if (!EX && a.count) {   // distances written back: whole 16-byte groups inside the bucket, single values at its ends
                            const u32 nst = inside[g] ? (ch ? 4u : 0u)
                                                      : (u32)(v[g].x != old[g].x) + (u32)(v[g].y != old[g].y) + (u32)(v[g].z != old[g].z) + (u32)(v[g].w != old[g].w);
                            const u32 tot = __reduce_add_sync(FULL, nst);
                            if (lane == 0 && tot) atomicAdd(reinterpret_cast<u64 *>(__cvta_shared_to_generic(cnt_s + 56)), (u64)tot);
                        }
                        // only distances that changed go back, as one predicated store per group; at a run's ends single values:
                        // never touch a neighbour bucket's positions, another warp may be writing them
                        float *pd = const_cast<float *>(base) + (size_t)dim * npad + g * 128;
                        stcg128_if(pd, v[g], ch && inside[g]);
                        if (hasedge) {
                            const bool e = ch && !inside[g];
                            stcg32_if(pd + 0, v[g].x, e && v[g].x != old[g].x);
                            stcg32_if(pd + 1, v[g].y, e && v[g].y != old[g].y);
                            stcg32_if(pd + 2, v[g].z, e && v[g].z != old[g].z);
                            stcg32_if(pd + 3, v[g].w, e && v[g].w != old[g].w);
                        }
                        // ascending positions: a lane keeps its first maximum
#define S_TRACK(E, OFF)                                              \
    if (v[g].E > best) {                                             \
        best = v[g].E;                                               \
        bi = p4 + OFF;                                               \
        _Pragma("unroll") for (int c = 0; c < DIM; ++c) bc[c] = x[c][g].E; \
    }
                        S_TRACK(x, 0)
                        S_TRACK(y, 1)
                        S_TRACK(z, 2)
                        S_TRACK(w, 3)
#undef S_TRACK
                    }
                }
                // this warp's maximum of the bucket and its lowest position (a warp without chunks reports "nothing")
                const float pv = fmaxf(best, 0.0f);
                const u32 m = __reduce_max_sync(FULL, __float_as_uint(pv));
                const u32 qpos = __reduce_min_sync(FULL, (__float_as_uint(pv) == m) ? bi : S_NONE);
                const u32 rec = pbuf + (fi * WPC + tw) * S_REC;
                // the record orders as ONE 64-bit key: (max bits, ~position) -- larger value first, then the lower position;
                // "nothing" is the smallest key there is
                if (qpos == S_NONE) {
                    if (lane == 0) sts128u(rec, 0u, 0u, 0u, 0u);
                } else if (bi == qpos) {   // positions are unique: exactly one lane
                    sts128u(rec, m, ~qpos, 0u, 0u);
                    s_sts128(rec + 16, bc[0], DIM > 1 ? bc[DIM > 1 ? 1 : 0] : 0.f, DIM > 2 ? bc[DIM > 2 ? 2 : 0] : 0.f, DIM > 3 ? bc[DIM > 3 ? 3 : 0] : 0.f);
                    if constexpr (DIM > 4)
                        s_sts128(rec + 32, bc[4], DIM > 5 ? bc[DIM > 5 ? 5 : 0] : 0.f, DIM > 6 ? bc[DIM > 6 ? 6 : 0] : 0.f, DIM > 7 ? bc[DIM > 7 ? 7 : 0] : 0.f);
                }
            }
            team_sync<WPC>(team);
            // ---- every warp merges the partial maxima into its own table: largest value, lowest position ------------------------
            for (u32 fi = 0; fi < nf; ++fi) {
                const u32 b = __shfl_sync(FULL, myb, fi);
                u64 bk = 0;
                u32 src = pbuf + (fi * WPC) * S_REC;
#pragma unroll
                for (int w = 0; w < WPC; ++w) {
                    const u32 rec = pbuf + (fi * WPC + w) * S_REC;
                    const uint4 pr = lds128u(rec);
                    const u64 key = (u64)pr.x << 32 | pr.y;
                    if (key > bk) bk = key, src = rec;
                }
                const u32 m = (u32)(bk >> 32), qpos = ~(u32)bk;   // (nothing at all: qpos = S_NONE)
                // (written as selects on purpose: an unrolled `if (slot == s) table[s] = ...` is turned into a dynamically indexed
                // store by the compiler, which moves the whole table to local memory)
                const bool mineb = (b & 31u) == lane;
#pragma unroll
                for (u32 s = 0; s < NW; ++s) {
                    const bool hit = mineb && (b >> 5) == s;
                    tmax[s] = hit ? __uint_as_float(m) : tmax[s];
                    tposn[s] = hit ? (u32)bk : tposn[s];
                }
                if (bk > fk) {   // (uniform over the warp)
                    fk = bk, fq = qpos;
                    if (lane == 0) {
                        const float4 c0v = s_lds128(src + 16);
                        s_sts128(fslot, c0v.x, c0v.y, c0v.z, c0v.w);
                        if constexpr (DIM > 4) {
                            const float4 c1v = s_lds128(src + 32);
                            s_sts128(fslot + 16, c1v.x, c1v.y, c1v.z, c1v.w);
                        }
                    }
                }
                if (mineb && (b >> 5) / BPL == tw) {   // the owner: its tests compare against the new maximum, its list is empty again
                    const float4 c0v = s_lds128(src + 16);
                    float4 c1v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if constexpr (DIM > 4) c1v = s_lds128(src + 32);
                    // the max point's coordinates for a LATER pick's arg-max (single writer; this pick's winner, if it is one of
                    // the buckets just passed over, is taken from fc above, so nobody reads this entry before the next barrier)
                    s_sts128(bmcs + b * PRB, c0v.x, c0v.y, c0v.z, c0v.w);
                    if constexpr (DIM > 4) s_sts128(bmcs + b * PRB + 16, c1v.x, c1v.y, c1v.z, c1v.w);
                    const float cc[8] = {c0v.x, c0v.y, c0v.z, c0v.w, c1v.x, c1v.y, c1v.z, c1v.w};
#pragma unroll
                    for (int j = 0; j < BPL; ++j) {
                        const bool hit = ((b >> 5) % BPL) == (u32)j;
                        if constexpr (WPC != 1) S_OWNMAX(j) = hit ? __uint_as_float(m) : S_OWNMAX(j);
                        np[j] = hit ? 0u : np[j];
#pragma unroll
                        for (int c = 0; c < DIM; ++c) bmc[j][c] = hit ? cc[c] : bmc[j][c];
                    }
                    sts32(brec + b * 16 + 8, 0u);
                }
            }
            bp ^= 1u;
        }

        // ---- 3. arg-max over all buckets, by every warp for itself: largest max, lowest position (KDLineTree.h:56-67) ---------
        u64 bestk = 0;
        u32 sb = 0;
#pragma unroll
        for (u32 s = 0; s < NW; ++s) {
            const u64 key = (u64)__float_as_uint(tmax[s]) << 32 | tposn[s];
            if (key > bestk) bestk = key, sb = s;
        }
        const u32 kmax = (u32)(bestk >> 32);
        const u32 M = __reduce_max_sync(FULL, kmax);
        const u32 minen = kmax == M ? (u32)bestk : 0u;
        const u32 curn = __reduce_max_sync(FULL, minen);   // the largest complement = the lowest position
        cur = ~curn;
        // positions are unique: exactly one lane holds the winner (redux 22 cycles; ballot + ffs + shfl ~90)
        const u32 bw = __reduce_max_sync(FULL, (kmax == M && (u32)bestk == curn) ? sb * 32 + lane : 0u);
        {   // the winner's point: a bucket passed over in this pick -> this warp's own slot (written by lane 0 above); an older
            // maximum -> the table its owner filled before an earlier barrier
            __syncwarp();
            const u32 from = cur == fq ? fslot : bmcs + bw * PRB;
            const float4 c0v = s_lds128(from);
            float4 c1v = make_float4(0.f, 0.f, 0.f, 0.f);
            if constexpr (DIM > 4) c1v = s_lds128(from + 16);
            const float cc[8] = {c0v.x, c0v.y, c0v.z, c0v.w, c1v.x, c1v.y, c1v.z, c1v.w};
#pragma unroll
            for (int c = 0; c < DIM; ++c) r[c] = cc[c];
        }
        // ---- output: positions are turned into original ids 32 picks at a time (wrapper.hpp:57-59) -----------------
        if (tw == 0) {
            const u32 *perm = reinterpret_cast<const u32 *>(rg) + (size_t)(dim + 1) * npad;
            u64 *out = a.out + (size_t)cloud * k;
            if ((t & 31u) == 0) idv = __ldg(perm + mypos);
            if ((t & 31u) == 1 && t > 1) out[t - 33 + lane] = (u64)idv;
            if (lane == (t & 31u)) mypos = cur;
        }
    }
    if (tw == 0) {
        const u32 *perm = reinterpret_cast<const u32 *>(rg) + (size_t)(dim + 1) * npad;
        u64 *out = a.out + (size_t)cloud * k;
        if (k > 32 && ((k - 1) & 31u) == 0) out[k - 33 + lane] = (u64)idv;   // the block whose lookup the last pick issued
        const u32 k0 = (k - 1) & ~31u;   // tail: picks [k0, k)
        if (k0 + lane < k) out[k0 + lane] = (u64)__ldg(perm + mypos);
    }
    if (!EX && a.count && lane == 0 && tw == 0) {
        u64 *cs = reinterpret_cast<u64 *>(__cvta_shared_to_generic(cnt_s));
        atomicAdd(cs + 4, (u64)(k - 1) * S);   // bucket tests
        atomicAdd(cs + 5, (u64)(k - 1));       // picks
        atomicAdd(cs + 6, 1ull);               // clouds
    }
    team_sync<WPC>(team);   // the team's shared memory is reused by its next cloud
#undef S_OWNMAX
}

template <int DIM, int WPC, int BPL, bool EX /* the cloud has exactly DIM dimensions: no per-component checks */>
__global__ void __launch_bounds__(S_THREADS, 1) kdline_stream_kernel(StreamArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ u32 sched[S_THREADS / 32];
    __shared__ u64 cnt[8];   // executed work of this CTA (a.count)
    // Left to itself the compiler re-reads the thread id (S2R, a 20-cycle scoreboard wait) and re-derives lane / team /
    // warp-in-team wherever it needs them -- 11 % of the instructions of the 4-warp-team kernel.  Read through an opaque
    // instruction it stays in ONE register; that register is only worth it where it does not turn into a spill (measured,
    // sampling ms without -> with: 512 clouds x 3-D on 4-warp teams 29.5 -> 27.5, 1024 x 3-D on 2-warp teams 39.2 -> 38.6,
    // 512 x 6-D 73.7 -> 77.9)
    u32 tid = threadIdx.x;
    if constexpr (DIM <= 4) asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
    const u32 warp = tid >> 5, lane = tid & 31u;
    const u32 team = warp / WPC, tw = warp % WPC;
    constexpr u32 teams = S_THREADS / 32 / WPC;
    if (threadIdx.x < 8) cnt[threadIdx.x] = 0;
    __syncthreads();
    // first wave: cloud = team * gridDim + cta (few clouds spread over SMs before they stack up on one); afterwards
    // clouds are handed out dynamically
    u32 cloud = team * gridDim.x + blockIdx.x;
    while (cloud < a.B) {
        stream_cloud<DIM, WPC, BPL, EX>(a, cloud, team, tw, lane, smem_u32(smem_raw) + team * a.team_bytes, smem_u32(cnt));
        if (tw == 0 && lane == 0) sched[team] = atomicAdd(a.counter, 1u) + teams * gridDim.x;
        team_sync<WPC>(team);
        cloud = sched[team];
    }
    if (!EX && a.count) {
        __syncthreads();
        if (threadIdx.x < 8) atomicAdd(&g_stream_wexec[threadIdx.x], cnt[threadIdx.x]);
    }
}

// ======================================================================================================
//  host side
// ======================================================================================================
static int stream_dim(int dim) { return dim <= 3 ? 3 : dim == 4 ? 4 : dim <= 6 ? 6 : 8; }

static size_t stream_team_bytes(int dimp, u32 wpc, u32 bpl, u32 R) {
    const size_t SP = 32u * wpc * bpl, PRB = (size_t)((dimp + 3) / 4) * 16, NW = wpc * bpl;
    const size_t PE = dimp <= 4 ? 16 : dimp <= 6 ? 24 : 32;   // PendEntry<DIM>::kBytes
    size_t b = (R * SP * PE + 15) & ~(size_t)15;
    b += SP * 16 + SP * PRB + ((NW + 3) & ~(size_t)3) * 4 + SP * 4;
    b += 2 * S_MAXF * wpc * S_REC + wpc * PRB;
    return (b + 15) & ~(size_t)15;
}

// one team size: buckets per lane, pending-list depth and shared memory of a CTA of S_THREADS / 32 / wpc teams
static bool stream_seg(int dimp, u32 S, u32 wpc, size_t clouds, int n_sms, StreamSeg *sg) {
    const size_t cap = 227 * 1024 - 256;
    const u32 bpl = wpc == 1 ? 4 : wpc == 2 ? 2 : (S <= 128 ? 1 : S <= 256 ? 2 : 4);
    if (32u * wpc * bpl < S) return false;
    const u32 teams = S_THREADS / 32 / wpc;
    u32 R = S_MAXR;
    while (R > 3 && teams * stream_team_bytes(dimp, wpc, bpl, R) > cap) --R;
    if (teams * stream_team_bytes(dimp, wpc, bpl, R) > cap) return false;
    sg->wpc = wpc;
    sg->bpl = bpl;
    sg->rs = R;
    sg->team_bytes = (u32)stream_team_bytes(dimp, wpc, bpl, R);
    sg->smem = (size_t)teams * sg->team_bytes;
    sg->clouds = (u32)clouds;
    sg->grid = (u32)(clouds < (size_t)n_sms ? clouds : (size_t)n_sms);   // spread over every SM before stacking teams
    return true;
}

bool plan_kdline_stream(size_t n, size_t dim, size_t h, size_t B, int n_sms, StreamPlan *pl) {
    if (dim == 0 || dim > 8 || h == 0 || h > 9 || n == 0 || B == 0) return false;
    const Tuning &tu = tuning();
    const u32 S = 1u << h;
    const int dimp = stream_dim((int)dim);
    // Warps per cloud (measured, 100 k-point clouds x 3, 2^7 buckets, one B200).  A pick is a dependent chain whose length hardly
    // depends on how many other teams share the SM: a cloud took ~85 / 44 / 28 ms on a team of 1 / 2 / 4 warps whether the SM
    // was full or not when the plan was made (35 / 21 ms on 2 / 4 warps with the final kernel: the same ratio), and an SM
    // holds 16 / 8 / 4 such teams.  So a batch wants FULL WAVES of narrow teams (most clouds per SM)
    // and a last, partial wave of wide teams (short chain) instead of a half-empty wave of narrow ones: 4096 clouds =
    // 3 x 1184 on two warps + 544 on four, not 3.46 waves of two-warp teams.  The cut is a small dynamic programme over the
    // remaining clouds in units of one wave of the widest team.  One-warp teams only have room for 4 pending samples per
    // bucket (+18 % DRAM traffic) and are opt-in (STREAM_SPLIT=2); records of more than 4 dimensions and more than 128
    // buckets need the shared memory / lanes of a 4-warp team.
    const u32 widths[3] = {1, 2, 4};
    const double cost[3] = {85.0, 44.0, 28.0};   // relative time of one wave per team size
    StreamSeg cand[3];
    bool ok[3];
    for (int i = 0; i < 3; ++i) ok[i] = stream_seg(dimp, S, widths[i], 1, n_sms, &cand[i]);
    if (!ok[2]) return false;
    const int split = tu.stream_split < 0 ? 1 : tu.stream_split;
    if (dimp > 4 || S > 128) ok[0] = ok[1] = false;
    if (split < 2) ok[0] = false;
    size_t take[3] = {0, 0, 0};   // clouds per team size
    if (tu.stream_warps == 1 || tu.stream_warps == 2 || tu.stream_warps == 4) {
        int i = tu.stream_warps == 1 ? 0 : tu.stream_warps == 2 ? 1 : 2;
        while (!stream_seg(dimp, S, widths[i], 1, n_sms, &cand[i])) ++i;   // a team too small for its shared memory: the next size up
        take[i] = B;
    } else if (split == 0) {
        const size_t slots = (size_t)(S_THREADS / 32) * n_sms;   // one-warp teams the GPU holds (2368)
        take[(ok[1] && B >= slots / 3) ? 1 : 2] = B;
    } else {
        const size_t unit = (size_t)(S_THREADS / 32 / 4) * n_sms;   // one wave of 4-warp teams (592)
        const size_t J = (B + unit - 1) / unit;
        std::vector<double> best(J + 1, 0.0);
        std::vector<int> pick(J + 1, 2);
        for (size_t j = 1; j <= J; ++j) {   // j units still to place
            best[j] = 1e300;
            for (int i = 0; i < 3; ++i) {
                if (!ok[i]) continue;
                const size_t wave = 4 / widths[i];   // units one wave of this team size takes
                const double c = cost[i] + (j > wave ? best[j - wave] : 0.0);
                if (c < best[j] - 1e-9) best[j] = c, pick[j] = i;
            }
        }
        size_t left = B;
        for (size_t j = J; j > 0 && left > 0;) {
            const int i = pick[j];
            const size_t wave = 4 / widths[i], cl = wave * unit < left ? wave * unit : left;
            take[i] += cl;
            left -= cl;
            j = j > wave ? j - wave : 0;
        }
    }
    pl->dimp = dimp;
    pl->nseg = 0;
    pl->desc[0] = 0;
    for (int i = 0; i < 3; ++i) {   // narrow teams first: the wide ones finish the batch
        if (!take[i]) continue;
        StreamSeg &sg = pl->seg[pl->nseg++];
        if (!stream_seg(dimp, S, widths[i], take[i], n_sms, &sg)) return false;
        const size_t len = strlen(pl->desc);
        snprintf(pl->desc + len, sizeof(pl->desc) - len, "%s%u clouds x WPC=%u (BPL=%u R=%u grid=%u smem=%zu)", len ? " + " : "", sg.clouds,
                 sg.wpc, sg.bpl, sg.rs, sg.grid, sg.smem);
    }
    return pl->nseg > 0;
}

template <int DIM, int WPC, int BPL, bool EX>
static cudaError_t launch_stream_t(const StreamSeg &pl, const StreamArgs &a, cudaStream_t st) {
    auto kern = kdline_stream_kernel<DIM, WPC, BPL, EX>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) return e;
    kern<<<pl.grid, S_THREADS, pl.smem, st>>>(a);
    return cudaGetLastError();
}

template <int DIM>
static cudaError_t launch_stream_d(const StreamSeg &pl, const StreamArgs &a, cudaStream_t st) {
    // the exact-dimension kernels exist for the shapes of BASELINE.json's configs[4] (3 and 6 dimensions, teams of 2 / 4 warps)
    if constexpr (DIM == 3 || DIM == 6) {
        if (a.dim == (u32)DIM && !a.count) {   // (the executed-work counters are compiled into the general kernels only)
            if (pl.wpc == 2) return launch_stream_t<DIM, 2, 2, true>(pl, a, st);
            if (pl.wpc == 4 && pl.bpl == 1) return launch_stream_t<DIM, 4, 1, true>(pl, a, st);
        }
    }
    if (pl.wpc == 1) return launch_stream_t<DIM, 1, 4, false>(pl, a, st);
    if (pl.wpc == 2) return launch_stream_t<DIM, 2, 2, false>(pl, a, st);
    if (pl.bpl == 1) return launch_stream_t<DIM, 4, 1, false>(pl, a, st);
    if (pl.bpl == 2) return launch_stream_t<DIM, 4, 2, false>(pl, a, st);
    return launch_stream_t<DIM, 4, 4, false>(pl, a, st);
}

cudaError_t stream_debug_counters(u64 *out16) { return cudaMemcpyFromSymbol(out16, g_stream_wexec, sizeof(u64) * 16); }

cudaError_t launch_kdline_stream(const StreamPlan &pl, unsigned char *region, size_t region_stride, const u64 *starts, u64 *out,
                                 u32 *counter, u32 B, u32 n, u32 dim, u32 k, u32 h, bool count, cudaStream_t st) {
    StreamArgs a;
    a.region_stride = region_stride;
    a.n = n;
    a.npad = (n + 31) & ~31u;
    a.dim = dim;
    a.k = k;
    a.S = 1u << h;
    a.nlo_pad = (a.S + 1 + 31) & ~31u;
    a.count = count ? 1u : 0u;
    a.prefetch = tuning().prefetch != 0 ? 1u : 0u;
    a.negzero = 0x8000000080000000ull;
    cudaError_t e = cudaMemsetAsync(counter, 0, 256, st);   // one hand-out counter per launch, 64 bytes apart
    if (e != cudaSuccess) return e;
    if (count) {
        void *sym = nullptr;
        if ((e = cudaGetSymbolAddress(&sym, g_stream_wexec)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(sym, 0, sizeof(u64) * 16, st)) != cudaSuccess) return e;
    }
    size_t b0 = 0;
    for (u32 i = 0; i < pl.nseg; ++i) {
        const StreamSeg &sg = pl.seg[i];
        a.region = region + b0 * region_stride;
        a.starts = starts ? starts + b0 : nullptr;
        a.out = out + b0 * (size_t)k;
        a.counter = counter + 16 * i;
        a.B = sg.clouds;
        a.R = sg.rs;
        a.team_bytes = sg.team_bytes;
        switch (pl.dimp) {
            case 3: e = launch_stream_d<3>(sg, a, st); break;
            case 4: e = launch_stream_d<4>(sg, a, st); break;
            case 6: e = launch_stream_d<6>(sg, a, st); break;
            default: e = launch_stream_d<8>(sg, a, st); break;
        }
        count_launch();
        if (e != cudaSuccess) return e;
        b0 += sg.clouds;
    }
    return b0 == B ? cudaSuccess : cudaErrorInvalidValue;
}

}  // namespace fps
