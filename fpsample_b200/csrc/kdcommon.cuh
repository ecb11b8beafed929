// kdcommon.cuh -- device helpers shared by the kd-line build kernels (kdline.cu: one CTA per cloud; kdbuild.cu:
// the whole grid per level).
#pragma once
#include "common.cuh"

namespace fps {

__device__ __forceinline__ u32 roundup32(u32 x) { return (x + 31u) & ~31u; }

// tight boxes of [s0,s1) split at sp: positions < sp go to boxL, the rest to boxR (ordered ints)
template <int DIM>
__device__ __forceinline__ void box_range(const float *q, u32 npad, u32 dim, u32 s0, u32 s1, u32 sp,
                                          int *boxL, int *boxR) {
    const u32 lane = lane_id();
    int lmin[DIM], lmax[DIM], rmin[DIM], rmax[DIM];
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
        lmin[c] = rmin[c] = 0x7fffffff;
        lmax[c] = rmax[c] = (int)0x80000000;
    }
    for (u32 i = s0 + lane; i < s1; i += 32) {
        const bool left = i < sp;
#pragma unroll
        for (int c = 0; c < DIM; ++c) {
            if (c < (int)dim) {
                int o = f2ord(q[(size_t)c * npad + i]);
                if (left) {
                    lmin[c] = min(lmin[c], o);
                    lmax[c] = max(lmax[c], o);
                } else {
                    rmin[c] = min(rmin[c], o);
                    rmax[c] = max(rmax[c], o);
                }
            }
        }
    }
    const bool anyL = s0 < min(s1, sp), anyR = max(s0, sp) < s1;
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
        if (c < (int)dim) {
            if (anyL) {
                int a = __reduce_min_sync(FULL, lmin[c]), b = __reduce_max_sync(FULL, lmax[c]);
                if (lane == 0) {
                    atomicMin(boxL + c, a);
                    atomicMax(boxL + dim + c, b);
                }
            }
            if (anyR) {
                int a = __reduce_min_sync(FULL, rmin[c]), b = __reduce_max_sync(FULL, rmax[c]);
                if (lane == 0) {
                    atomicMin(boxR + c, a);
                    atomicMax(boxR + dim + c, b);
                }
            }
        }
    }
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

}  // namespace fps
