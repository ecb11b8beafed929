// npdu.cu -- fps_npdu_sampling: FPS with the reference's "nearest point distance updating"
// heuristic over an INDEX window, src/lib.cpp:272-340.  Not exact FPS: after one full min-update against the start point,
// a pick only min-updates the points whose index lies within k/2 of it (the window is shifted, not shrunk, at the array
// ends, lib.cpp:296-300), then the arg-max runs over ALL points with strict '>' from -1 (lowest index among equal maxima,
// lib.cpp:311-315).  The reference pays O(n) per pick for that arg-max; here one CTA per cloud keeps the maximum of every
// 256-point segment as a 64-bit key (distance bits, ~index) in shared memory, so a pick costs the window (w + 1 distances),
// the few segments the window touches, and a reduction over n / 256 keys.  Distances in the reference's arithmetic
// (individually rounded sub / mul / add in dimension order, common.cuh); the running distances live in shared memory for
// clouds of up to 16 384 points, in global memory (L2) beyond.
#include "common.cuh"
#include "engine.h"

namespace fps {

constexpr u32 NP_T = 256, NP_NW = NP_T / 32, NP_SEG = 256, NP_MAXDIM = 64;

struct NpduArgs {
    const float *pts;    // [B][n][dim]
    float *dm;           // [B][npad] running distances
    const u64 *starts;   // nullptr or [B]
    u64 *out;            // [B][k]
    u32 B, n, npad, dim, k, w;
    u32 dm_smem;         // the running distances fit shared memory next to the segment keys (n <= 16 384)
};

__device__ __forceinline__ float np_sqdist(const float *p, const float *r, u32 dim) {
    float t = __fsub_rn(p[0], r[0]);
    float acc = __fmul_rn(t, t);
    for (u32 j = 1; j < dim; ++j) {
        t = __fsub_rn(p[j], r[j]);
        acc = __fadd_rn(acc, __fmul_rn(t, t));
    }
    return acc;
}

__global__ void __launch_bounds__(NP_T) npdu_kernel(NpduArgs a) {
    extern __shared__ __align__(8) u64 segkey[];   // [ceil(n / NP_SEG)]
    __shared__ u64 wred[NP_NW];
    __shared__ float sref[NP_MAXDIM];
    const u32 tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    const u32 n = a.n, dim = a.dim, nseg = (n + NP_SEG - 1) / NP_SEG;
    for (u32 cloud = blockIdx.x; cloud < a.B; cloud += gridDim.x) {
        const float *p = a.pts + (size_t)cloud * n * dim;
        float *dm = a.dm_smem ? reinterpret_cast<float *>(segkey + nseg) : a.dm + (size_t)cloud * a.npad;
        u64 *out = a.out + (size_t)cloud * a.k;
        u32 cur = a.starts ? (u32)a.starts[cloud] : 0u;
        __syncthreads();
        if (tid < dim) sref[tid] = p[(size_t)cur * dim + tid];
        __syncthreads();
        for (u32 i = tid; i < n; i += NP_T) dm[i] = np_sqdist(p + (size_t)i * dim, sref, dim);   // min(+inf, d), lib.cpp:319-326
        __syncthreads();
        auto seg_key = [&](u32 sg) {   // one warp: the segment's largest distance, lowest index among equals
            u64 best = 0;
            const u32 i1 = min(n, (sg + 1) * NP_SEG);
            for (u32 i = sg * NP_SEG + lane; i < i1; i += 32) {
                const u64 key = make_key(dm[i], ~i);
                best = key > best ? key : best;
            }
            best = warp_max_key(best);
            if (lane == 0) segkey[sg] = best;
        };
        for (u32 sg = warp; sg < nseg; sg += NP_NW) seg_key(sg);
        if (tid == 0) out[0] = cur;
        const long long P = (long long)n, hw = (long long)(a.w / 2);
        for (u32 t = 1; t < a.k; ++t) {
            long long s = (long long)cur - hw, e = (long long)cur + hw;   // lib.cpp:294-300
            if (s < 0) e -= s, s = 0;
            if (e >= P) {
                s = s - (e - P + 1);
                if (s < 0) s = 0;
                e = P - 1;
            }
            for (long long i = s + tid; i <= e; i += NP_T) {
                const float v = np_sqdist(p + (size_t)i * dim, sref, dim);
                if (v < dm[i]) dm[i] = v;
            }
            __syncthreads();
            for (u32 sg = (u32)(s / NP_SEG) + warp; sg <= (u32)(e / NP_SEG); sg += NP_NW) seg_key(sg);
            __syncthreads();
            u64 best = 0;
            for (u32 sg = tid; sg < nseg; sg += NP_T) best = segkey[sg] > best ? segkey[sg] : best;
            best = warp_max_key(best);
            if (lane == 0) wred[warp] = best;
            __syncthreads();
            best = wred[0];
#pragma unroll
            for (u32 wv = 1; wv < NP_NW; ++wv) best = wred[wv] > best ? wred[wv] : best;
            cur = ~(u32)best;
            if (tid == 0) out[t] = cur;
            if (tid < dim) sref[tid] = p[(size_t)cur * dim + tid];
            __syncthreads();
        }
    }
}

// ======================================================================================================
//  fps_npdu_kdtree_sampling (second half): src/lib.cpp:369-465.  After a full min-update against the
//  start point, a pick only min-updates its k NEAREST points (the reference asks nanoflann for them, lib.cpp:421-436), then
//  the arg-max runs over ALL points with strict '>' from -1 (lowest index among equal maxima, lib.cpp:438-442).  What the
//  reference computes is therefore a function of the SET of the k nearest points in its own binary32 distance
//  (dim-order sub / mul / add, lib.cpp:33-41) -- verified against the compiled reference (tests/golden, oracle) -- so no kd
//  tree is needed here: one CTA per cloud computes the n distances to the pick, finds the k-th smallest with a 4-pass radix
//  select over the float bits (shared-memory histograms; the first pass rides on the distance pass), min-updates everything
//  below it and recomputes the arg-max in the same pass.  Ties AT the k-th distance (more equal candidates than places) are
//  taken in index order; the reference takes them in nanoflann's traversal order, so on clouds with exact distance ties the
//  two may pick different (equally near) neighbours -- the one documented difference.
constexpr u32 NK_T = 512, NK_NW = NK_T / 32;

struct NpduKnnArgs {
    const float *pts;    // [B][n][dim]
    float *dm;           // [B][npad] running distances (global copy, used when they do not fit shared memory)
    float *dq;           // [B][npad] distances to the current pick
    const u64 *starts;
    u64 *out;
    u32 B, n, npad, dim, k, w;
    u32 in_smem;
};

__global__ void __launch_bounds__(NK_T) npdu_knn_kernel(NpduKnnArgs a) {
    extern __shared__ __align__(16) float nk_sm[];   // [dm n][dq n] when in_smem
    __shared__ u32 hist[256];
    __shared__ u64 wred[NK_NW];
    __shared__ float sref[NP_MAXDIM];
    __shared__ u32 sel[3];   // prefix, remaining, count in the chosen bin
    const u32 tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    const u32 n = a.n, dim = a.dim;
    const u32 kk = a.w < n ? a.w : n;   // k_use, lib.cpp:407
    for (u32 cloud = blockIdx.x; cloud < a.B; cloud += gridDim.x) {
        const float *p = a.pts + (size_t)cloud * n * dim;
        float *dm = a.in_smem ? nk_sm : a.dm + (size_t)cloud * a.npad;
        u32 *dq = reinterpret_cast<u32 *>(a.in_smem ? nk_sm + n : a.dq + (size_t)cloud * a.npad);
        u64 *out = a.out + (size_t)cloud * a.k;
        u32 cur = a.starts ? (u32)a.starts[cloud] : 0u;
        __syncthreads();
        if (tid < dim) sref[tid] = p[(size_t)cur * dim + tid];
        __syncthreads();
        for (u32 i = tid; i < n; i += NK_T) dm[i] = np_sqdist(p + (size_t)i * dim, sref, dim);   // min(+inf, d), lib.cpp:446-453
        if (tid == 0) out[0] = cur;
        __syncthreads();
        for (u32 t = 1; t < a.k; ++t) {
            u32 tau = 0xffffffffu, need = 0, cnt = 0;
            if (t > 1 && kk < n) {   // (the k nearest of the START point change nothing: they were all updated above)
                // ---- distances to the pick + histogram of their top byte --------------------------------------------------------
                if (tid < 256) hist[tid] = 0;
                __syncthreads();
                for (u32 i = tid; i < n; i += NK_T) {
                    const u32 b = __float_as_uint(np_sqdist(p + (size_t)i * dim, sref, dim));
                    dq[i] = b;
                    atomicAdd(&hist[b >> 24], 1u);
                }
                __syncthreads();
                // ---- radix select of the kk-th smallest (distances are >= 0: their bits order like the values) ---------------------
                u32 prefix = 0, remaining = kk;
                for (int shift = 24; shift >= 0; shift -= 8) {
                    if (warp == 0) {   // the bin holding the remaining-th element: 8 bins per lane, a warp scan, then the lane's own 8
                        u32 h[8], s = 0;
#pragma unroll
                        for (int x = 0; x < 8; ++x) h[x] = hist[lane * 8 + x], s += h[x];
                        u32 inc = s;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const u32 v = __shfl_up_sync(FULL, inc, o);
                            if ((int)lane >= o) inc += v;
                        }
                        const u32 first = __ffs(__ballot_sync(FULL, inc >= remaining)) - 1;
                        if (lane == first) {
                            u32 before = inc - s, bin = 0, c = 0;
#pragma unroll
                            for (int x = 0; x < 8; ++x)
                                if (c == 0) {
                                    if (before + h[x] >= remaining) bin = x, c = h[x];
                                    else before += h[x];
                                }
                            sel[0] = prefix | ((lane * 8 + bin) << shift);
                            sel[1] = remaining - before;
                            sel[2] = c;
                        }
                    }
                    __syncthreads();
                    prefix = sel[0], remaining = sel[1], cnt = sel[2];
                    if (shift == 0) break;
                    if (tid < 256) hist[tid] = 0;
                    __syncthreads();
                    const u32 hi_mask = 0xffffffffu << shift;
                    for (u32 i = tid; i < n; i += NK_T) {
                        const u32 b = dq[i];
                        if ((b & hi_mask) == prefix) atomicAdd(&hist[(b >> (shift - 8)) & 255u], 1u);
                    }
                    __syncthreads();
                }
                tau = prefix, need = remaining;   // `need` of the `cnt` points at distance tau belong to the k nearest
                if (need < cnt) {   // more equal candidates than places: index order (rare; a serial walk by one thread)
                    if (tid == 0) {
                        u32 left = cnt - need;   // the LAST `left` of them (by index) are dropped: moved out of reach
                        for (u32 i = n; i-- > 0 && left;)
                            if (dq[i] == tau) dq[i] = 0xffffffffu, --left;
                    }
                    __syncthreads();
                }
            }
            // ---- min-update of the k nearest (all points when kk == n), arg-max over everything in the same pass ------------------
            u64 best = 0;
            const bool upd = t > 1;
            for (u32 i = tid; i < n; i += NK_T) {
                float v = dm[i];
                if (upd) {
                    if (kk < n) {
                        const u32 b = dq[i];
                        if (b <= tau) {
                            const float d = __uint_as_float(b);
                            if (d < v) dm[i] = v = d;
                        }
                    } else {
                        const float d = np_sqdist(p + (size_t)i * dim, sref, dim);
                        if (d < v) dm[i] = v = d;
                    }
                }
                const u64 key = make_key(v, ~i);
                best = key > best ? key : best;
            }
            best = warp_max_key(best);
            if (lane == 0) wred[warp] = best;
            __syncthreads();
            best = wred[0];
#pragma unroll
            for (u32 wv = 1; wv < NK_NW; ++wv) best = wred[wv] > best ? wred[wv] : best;
            cur = ~(u32)best;
            if (tid == 0) out[t] = cur;
            if (tid < dim) sref[tid] = p[(size_t)cur * dim + tid];
            __syncthreads();
        }
    }
}

size_t npdu_knn_workspace_bytes(size_t B, size_t n) { return 2 * B * ((n + 31) & ~(size_t)31) * sizeof(float) + 256; }

cudaError_t launch_npdu_knn(const float *pts, size_t B, size_t n, size_t dim, size_t k, size_t w, const u64 *starts, u64 *out,
                            void *ws, int n_sms, cudaStream_t st) {
    if (dim == 0 || dim > NP_MAXDIM) return cudaErrorNotSupported;
    const size_t npad = (n + 31) & ~(size_t)31;
    const bool in_smem = 2 * n * sizeof(float) <= 160 * 1024;
    const size_t smem = in_smem ? 2 * n * sizeof(float) : 0;
    NpduKnnArgs a;
    a.pts = pts;
    a.dm = static_cast<float *>(ws);
    a.dq = a.dm + B * npad;
    a.starts = starts;
    a.out = out;
    a.B = (u32)B, a.n = (u32)n, a.npad = (u32)npad, a.dim = (u32)dim, a.k = (u32)k;
    a.w = (u32)(w > 0xffffffffull ? 0xffffffffull : w);
    a.in_smem = in_smem ? 1u : 0u;
    cudaError_t e = cudaFuncSetAttribute(npdu_knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    size_t grid = (size_t)n_sms * (smem > 100 * 1024 ? 1 : smem > 48 * 1024 ? 2 : 4);
    if (grid > B) grid = B;
    npdu_knn_kernel<<<(unsigned)grid, NK_T, smem, st>>>(a);
    count_launch();
    return cudaGetLastError();
}

size_t npdu_workspace_bytes(size_t B, size_t n) { return B * ((n + 31) & ~(size_t)31) * sizeof(float) + 256; }

cudaError_t launch_npdu(const float *pts, size_t B, size_t n, size_t dim, size_t k, size_t w, const u64 *starts, u64 *out,
                        void *ws, int n_sms, cudaStream_t st) {
    if (dim == 0 || dim > NP_MAXDIM) return cudaErrorNotSupported;
    size_t smem = ((n + NP_SEG - 1) / NP_SEG) * sizeof(u64);
    if (smem > 200 * 1024) return cudaErrorNotSupported;
    const bool dm_smem = n <= 16384;   // 64 KB of distances: still three or more clouds per SM
    if (dm_smem) smem += n * sizeof(float);
    NpduArgs a;
    a.pts = pts;
    a.dm = static_cast<float *>(ws);
    a.starts = starts;
    a.out = out;
    a.B = (u32)B, a.n = (u32)n, a.npad = (u32)((n + 31) & ~(size_t)31), a.dim = (u32)dim, a.k = (u32)k;
    a.w = (u32)(w > 0xffffffffull ? 0xffffffffull : w);
    a.dm_smem = dm_smem ? 1u : 0u;
    cudaError_t e = cudaFuncSetAttribute(npdu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    size_t grid = (size_t)n_sms * 8;
    if (grid > B) grid = B;
    npdu_kernel<<<(unsigned)grid, NP_T, smem, st>>>(a);
    count_launch();
    return cudaGetLastError();
}

}  // namespace fps
