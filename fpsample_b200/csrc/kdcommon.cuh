// kdcommon.cuh -- device helpers shared by the kd-line build kernels (kdline.cu: one CTA per cloud; kdbuild.cu:
// the whole grid per level).
#pragma once
#include "common.cuh"
#include "seqsum.cuh"

namespace fps {

__device__ __forceinline__ u32 roundup32(u32 x) { return (x + 31u) & ~31u; }

// tight box of the positions [s0, s1) (one warp), folded into `box` (ordered ints: lows then highs) with atomics
template <int DIM>
__device__ __forceinline__ void box_fold(const float *q, u32 npad, u32 dim, u32 s0, u32 s1, int *box) {
    if (s0 >= s1) return;   // warp-uniform
    const u32 lane = lane_id();
    float mn[DIM], mx[DIM];
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
        mn[c] = __int_as_float(0x7f800000);    // +inf
        mx[c] = __int_as_float(0xff800000);    // -inf
    }
    for (u32 i = s0 + lane; i < s1; i += 32) {
#pragma unroll
        for (int c = 0; c < DIM; ++c) {
            if (c < (int)dim) {
                const float v = q[(size_t)c * npad + i];
                mn[c] = fminf(mn[c], v);
                mx[c] = fmaxf(mx[c], v);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
        if (c < (int)dim) {
            // +-0 compare equal but order differently as ints; either is a correct bound (the box only enters
            // subtractions and comparisons), lanes without points contribute +-inf
            const int a = __reduce_min_sync(FULL, f2ord(mn[c])), b = __reduce_max_sync(FULL, f2ord(mx[c]));
            if (lane == 0) {
                atomicMin(box + c, a);
                atomicMax(box + dim + c, b);
            }
        }
    }
}

// tight boxes of [s0,s1) split at sp: positions < sp go to boxL, the rest to boxR (ordered ints)
template <int DIM>
__device__ __forceinline__ void box_range(const float *q, u32 npad, u32 dim, u32 s0, u32 s1, u32 sp,
                                          int *boxL, int *boxR) {
    box_fold<DIM>(q, npad, dim, s0, min(s1, sp), boxL);
    box_fold<DIM>(q, npad, dim, max(s0, sp), s1, boxR);
}

// strictly sequential binary32 sum of src[0..count) in order (KDTreeBase.h:151-158), computed redundantly by all 32
// lanes: lanes fetch 128 consecutive values (4 coalesced loads), the add chain then consumes them through shuffles.
// The next 128 values are in flight while the chain runs: the dependent FADD chain (4 cycles per element) is the only
// thing on the critical path.
__device__ __forceinline__ float seq_sum(const float *src, u32 count) {
    const u32 lane = lane_id();
    float sum = 0.0f;
    u32 i = 0;
    float x0 = 0.f, x1 = 0.f, x2 = 0.f, x3 = 0.f;
    if (count >= 128) {
        x0 = src[lane];
        x1 = src[32 + lane];
        x2 = src[64 + lane];
        x3 = src[96 + lane];
    }
    for (; i + 128 <= count; i += 128) {
        float n0 = 0.f, n1 = 0.f, n2 = 0.f, n3 = 0.f;
        if (i + 256 <= count) {
            n0 = src[i + 128 + lane];
            n1 = src[i + 160 + lane];
            n2 = src[i + 192 + lane];
            n3 = src[i + 224 + lane];
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) sum = __fadd_rn(sum, __shfl_sync(FULL, x0, j));
#pragma unroll
        for (int j = 0; j < 32; ++j) sum = __fadd_rn(sum, __shfl_sync(FULL, x1, j));
#pragma unroll
        for (int j = 0; j < 32; ++j) sum = __fadd_rn(sum, __shfl_sync(FULL, x2, j));
#pragma unroll
        for (int j = 0; j < 32; ++j) sum = __fadd_rn(sum, __shfl_sync(FULL, x3, j));
        x0 = n0;
        x1 = n1;
        x2 = n2;
        x3 = n3;
    }
    for (; i < count; i += 32) {
        float x = (i + lane < count) ? src[i + lane] : 0.0f;
        const u32 m = min(32u, count - i);
        for (u32 j = 0; j < m; ++j) sum = __fadd_rn(sum, __shfl_sync(FULL, x, j));
    }
    return sum;
}

// Same strictly sequential sum, fed through a warp-private shared-memory staging buffer (2 x 128 floats) instead of
// shuffles: per 128 elements the warp does 4 coalesced loads + 4 stores, then 32 broadcast 128-bit shared loads feed
// the dependent FADD chain, while the next 128 values are already in flight.  Shuffle throughput, not the 4-cycle
// FADD latency, bounded the shuffle version (measured 10 cycles per element on B200).
__device__ __forceinline__ float seq_sum_staged(const float *src, u32 count, float *buf /* [256], warp-private */) {
    const u32 lane = lane_id();
    float sum = 0.0f;
    u32 i = 0;
    if (count >= 128) {
        float x0 = src[lane], x1 = src[32 + lane], x2 = src[64 + lane], x3 = src[96 + lane];
        u32 cur = 0;
        buf[lane] = x0;
        buf[32 + lane] = x1;
        buf[64 + lane] = x2;
        buf[96 + lane] = x3;
        __syncwarp();
        for (; i + 128 <= count; i += 128) {
            const bool more = i + 256 <= count;
            if (more) {
                x0 = src[i + 128 + lane];
                x1 = src[i + 160 + lane];
                x2 = src[i + 192 + lane];
                x3 = src[i + 224 + lane];
            }
            const float4 *b4 = reinterpret_cast<const float4 *>(buf + cur * 128);
            // software pipeline: the 8 broadcast loads of the next 32 values are issued before the 32 dependent adds
            float4 va[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) va[j] = b4[j];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                float4 vb[8];
                if (g < 3) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) vb[j] = b4[8 * (g + 1) + j];
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    sum = __fadd_rn(sum, va[j].x);
                    sum = __fadd_rn(sum, va[j].y);
                    sum = __fadd_rn(sum, va[j].z);
                    sum = __fadd_rn(sum, va[j].w);
                }
                if (g < 3) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) va[j] = vb[j];
                }
            }
            cur ^= 1;
            if (more) {
                float *nb = buf + cur * 128;
                nb[lane] = x0;
                nb[32 + lane] = x1;
                nb[64 + lane] = x2;
                nb[96 + lane] = x3;
            }
            __syncwarp();
        }
    }
    for (; i < count; i += 32) {
        float x = (i + lane < count) ? src[i + lane] : 0.0f;
        const u32 m = min(32u, count - i);
        for (u32 j = 0; j < m; ++j) sum = __fadd_rn(sum, __shfl_sync(FULL, x, j));
    }
    return sum;
}

// The same strictly sequential sum when the column already sits in shared memory: broadcast 128-bit loads feed the
// dependent FADD chain directly (every lane computes the same sum), the next 32 values are loaded while the
// current 32 are added -- the 4-cycle FADD latency is the only thing on the critical path.
__device__ __forceinline__ float seq_sum_smem(const float *src, u32 count) {
    float sum = 0.0f;
    u32 i = 0;
    while (i < count && (smem_u32(src + i) & 15u)) sum = __fadd_rn(sum, src[i++]);   // up to 3 values
    const float4 *p = reinterpret_cast<const float4 *>(src + i);
    const u32 nblk = (count - i) >> 4;   // blocks of 16 values = 4 float4 (the kernel runs under a 64-register cap)
    if (nblk) {
        float4 a0 = p[0], a1 = p[1], a2 = p[2], a3 = p[3];
        for (u32 b = 1; b <= nblk; ++b) {
            const u32 nb = b < nblk ? b : b - 1;   // last round reloads its own block: no branch in the chain
            const float4 n0 = p[4 * nb], n1 = p[4 * nb + 1], n2 = p[4 * nb + 2], n3 = p[4 * nb + 3];
            sum = __fadd_rn(sum, a0.x);
            sum = __fadd_rn(sum, a0.y);
            sum = __fadd_rn(sum, a0.z);
            sum = __fadd_rn(sum, a0.w);
            sum = __fadd_rn(sum, a1.x);
            sum = __fadd_rn(sum, a1.y);
            sum = __fadd_rn(sum, a1.z);
            sum = __fadd_rn(sum, a1.w);
            sum = __fadd_rn(sum, a2.x);
            sum = __fadd_rn(sum, a2.y);
            sum = __fadd_rn(sum, a2.z);
            sum = __fadd_rn(sum, a2.w);
            sum = __fadd_rn(sum, a3.x);
            sum = __fadd_rn(sum, a3.y);
            sum = __fadd_rn(sum, a3.z);
            sum = __fadd_rn(sum, a3.w);
            a0 = n0, a1 = n1, a2 = n2, a3 = n3;
        }
        i += nblk << 4;
    }
    for (; i < count; ++i) sum = __fadd_rn(sum, src[i]);
    return sum;
}

// The same strictly sequential sum over a column in GLOBAL memory, fed by TMA: lane 0 keeps three 2 KB tiles in
// flight (cp.async.bulk global -> this warp's private shared-memory ring, completion counted on an mbarrier per
// stage) while all lanes run the dependent FADD chain over the tile that has landed, straight from broadcast 128-bit
// shared loads.  One warp is its own producer and consumer; L2 / HBM latency is off the chain entirely.
// ring: 4 x 512 floats, 16-byte aligned, warp-private; bars: 4 mbarriers initialised to count 1; phase: per-stage
// parity bits carried across calls.
constexpr u32 SS_TILE = 512, SS_STAGES = 4;
__device__ __forceinline__ float seq_sum_tma(const float *src, u32 count, float *ring, u64 *bars, u32 &phase) {
    const u32 lane = lane_id();
    float sum = 0.0f;
    u32 i = 0;
    while (i < count && (reinterpret_cast<uintptr_t>(src + i) & 15u)) sum = __fadd_rn(sum, src[i++]);   // up to 3 values
    const float *tsrc = src + i;
    const u32 ntile = (count - i) / SS_TILE;
    if (ntile) {
        if (lane == 0) {
            for (u32 s = 0; s < SS_STAGES - 1 && s < ntile; ++s) {
                mbar_arrive_expect_tx(smem_u32(&bars[s]), SS_TILE * 4);
                tma_bulk_g2s(smem_u32(ring + s * SS_TILE), tsrc + (size_t)s * SS_TILE, SS_TILE * 4, smem_u32(&bars[s]));
            }
        }
        u32 tie_hint = 0;
        for (u32 t = 0; t < ntile; ++t) {
            const u32 st = t % SS_STAGES;
            __syncwarp();   // everybody is done with the stage that is refilled next
            if (lane == 0 && t + SS_STAGES - 1 < ntile) {
                const u32 s2 = (t + SS_STAGES - 1) % SS_STAGES;
                mbar_arrive_expect_tx(smem_u32(&bars[s2]), SS_TILE * 4);
                tma_bulk_g2s(smem_u32(ring + s2 * SS_TILE), tsrc + (size_t)(t + SS_STAGES - 1) * SS_TILE, SS_TILE * 4,
                             smem_u32(&bars[s2]));
            }
            mbar_wait_cluster(smem_u32(&bars[st]), (phase >> st) & 1u);
            phase ^= 1u << st;
            // the tile as one integer prefix scan where the running sum stays inside its binade (seqsum.cuh), else the chain
            if (seq_sum_tile<SS_TILE / 32>(smem_u32(ring + st * SS_TILE), sum, tie_hint)) continue;
            const float4 *p = reinterpret_cast<const float4 *>(ring + st * SS_TILE);
            float4 a[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = p[j];
#pragma unroll 1
            for (u32 b = 1; b <= SS_TILE / 32; ++b) {
                const u32 nb = b < SS_TILE / 32 ? b : b - 1;   // the last round reloads its own block: no branch in the chain
                float4 nx[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) nx[j] = p[8 * nb + j];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    sum = __fadd_rn(sum, a[j].x);
                    sum = __fadd_rn(sum, a[j].y);
                    sum = __fadd_rn(sum, a[j].z);
                    sum = __fadd_rn(sum, a[j].w);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] = nx[j];
            }
        }
        i += ntile * SS_TILE;
    }
    // tail (< 512 values): lanes fetch 32 at a time, the chain consumes them through shuffles
    for (; i < count; i += 32) {
        const float x = (i + lane < count) ? src[i + lane] : 0.0f;
        const u32 m = min(32u, count - i);
        for (u32 j = 0; j < m; ++j) sum = __fadd_rn(sum, __shfl_sync(FULL, x, j));
    }
    return sum;
}

}  // namespace fps
