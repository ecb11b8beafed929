// kdtree.cu -- the FULL kd tree of `bucket_fps_kdtree_sampling`: reference
// src/_ext/KDTree.h:13-52 (leaf rule count == 1, :27), src/wrapper.hpp:29-43,101-116, src/lib.cpp:467-520.
//
// What the reference computes (verified against the compiled reference on uniform, lidar-like and tie-lattice clouds,
// tests/test_oracle.py::test_kdtree_is_vanilla_over_the_full_permutation): the tree build permutes the point array
// exactly like the kd-line build (src/_ext/KDTreeBase.h:84-207: first-max-span split dimension, split value =
// SEQUENTIAL binary32 sum / count, in-place Hoare partition) but down to single points; sampling starts at the point
// sitting at POSITION start_idx after that permutation (wrapper.hpp:36-37); every interior node keeps the larger of its
// children's maxima, the RIGHT child winning ties (KDNode.h:41-46), so a tie goes to the HIGHEST position; the lazy
// delay lists (KDNode.h:120-166) never change a distance that the eager recurrence would not.  The result is therefore
// exact FPS over the fully permuted array with vanilla's own tie rule -- this file builds the permutation on the GPU and
// writes the permuted rows; the vanilla kernels (vanilla.cu) sample them; kdtree_map_kernel turns positions into ids.
//
// Build: one CTA per cloud, level-synchronous over a list of open segments [lo, hi) (count >= 2, not all points equal),
// one WARP per segment: tight box -> split dimension -> sequential mean (the dependent FADD chain, kdcommon.cuh) ->
// count -> the Hoare loop in closed form (k-th misplaced element from the left swaps with the k-th from the right) ->
// two child segments.  A segment whose box has zero extent in every dimension is closed at once: all its points are
// equal, the reference's partition clamps (KDTreeBase.h:142-146) and never moves them again.
#include <cfloat>

#include "common.cuh"
#include "engine.h"
#include "kdcommon.cuh"

namespace fps {

constexpr u32 KT_T = 1024;   // threads per CTA (32 warps = 32 segments in flight per cloud)

struct KdtreeArgs {
    const float *pts;        // [B][n][dim] row-major
    unsigned char *region;   // per cloud: [q dim*npad f32][scr npad u32][perm npad u32][listA npad u32][listB npad u32]
    size_t region_stride;
    float *rows;             // out [B][n][dim]: the permuted cloud, row-major (input of the vanilla kernels)
    u32 B, n, npad, dim;
};

template <int DIM>
__global__ void __launch_bounds__(KT_T, 1) kdtree_build_kernel(KdtreeArgs a) {
    __shared__ u32 s_cnt[2];
    __shared__ __align__(16) float s_chain[32 * 256];   // per-warp staging of the sequential sum
    const u32 tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    const u32 n = a.n, npad = a.npad, dim = a.dim;
    for (u32 cloud = blockIdx.x; cloud < a.B; cloud += gridDim.x) {
        unsigned char *rg = a.region + (size_t)cloud * a.region_stride;
        float *q = reinterpret_cast<float *>(rg);
        u32 *scr = reinterpret_cast<u32 *>(rg) + (size_t)dim * npad;
        u32 *perm = scr + npad;
        u32 *list[2] = {perm + npad, perm + 2 * (size_t)npad};
        const float *g = a.pts + (size_t)cloud * n * dim;
        for (u32 f = tid; f < n * dim; f += KT_T) {
            const u32 i = f / dim, c = f - i * dim;
            q[(size_t)c * npad + i] = g[f];
        }
        for (u32 i = tid; i < n; i += KT_T) perm[i] = i;
        if (tid == 0) {
            s_cnt[0] = n >= 2 ? 1u : 0u;
            s_cnt[1] = 0;
            list[0][0] = 0;
            list[0][1] = n;
        }
        __syncthreads();
        u32 cur = 0;
        for (;;) {
            const u32 nseg = s_cnt[cur];
            if (nseg == 0) break;
            for (u32 sgi = warp; sgi < nseg; sgi += KT_T / 32) {
                const u32 lo = list[cur][2 * sgi], hi = list[cur][2 * sgi + 1], count = hi - lo;
                // tight box, split dimension = first dimension of strictly largest extent (KDTreeBase.h:160-179)
                float mn[DIM], mx[DIM];
#pragma unroll
                for (int c = 0; c < DIM; ++c) {
                    mn[c] = __int_as_float(0x7f800000);
                    mx[c] = __int_as_float(0xff800000);
                }
                for (u32 i = lo + lane; i < hi; i += 32) {
#pragma unroll
                    for (int c = 0; c < DIM; ++c)
                        if (c < (int)dim) {
                            const float v = q[(size_t)c * npad + i];
                            mn[c] = fminf(mn[c], v);
                            mx[c] = fmaxf(mx[c], v);
                        }
                }
                u32 sd = 0;
                float span = 0.0f;
#pragma unroll
                for (int c = 0; c < DIM; ++c)
                    if (c < (int)dim) {
                        const float l = ord2f(__reduce_min_sync(FULL, f2ord(mn[c]))), h2 = ord2f(__reduce_max_sync(FULL, f2ord(mx[c])));
                        const float s = __fsub_rn(h2, l);
                        if (s > span) {
                            span = s;
                            sd = (u32)c;
                        }
                    }
                if (!(span > 0.0f)) continue;   // every point of the segment is the same point: never permuted again
                float *col = q + (size_t)sd * npad;
                const float sum = seq_sum_staged(col + lo, count, s_chain + warp * 256);   // KDTreeBase.h:151-158
                const float val = __fdiv_rn(sum, __uint2float_rn(count));
                u32 m = 0;
                for (u32 i = lo + lane; i < hi; i += 32) m += (col[i] < val) ? 1u : 0u;
                m = __reduce_add_sync(FULL, m);
                u32 lim = m;
                if (m == 0) lim = 1;
                else if (m == count) lim = count - 1;
                else {
                    // the Hoare loop (KDTreeBase.h:123-149) in closed form: misplaced elements of the left part in
                    // ascending order pair with misplaced elements of the right part in descending order
                    u32 base = 0;
                    for (u32 i0 = lo; i0 < hi; i0 += 32) {
                        const u32 i = i0 + lane;
                        const bool in = i < hi;
                        const bool f = in && (col[i] < val);
                        const u32 mask = __ballot_sync(FULL, f);
                        const u32 pre = base + __popc(mask & ((1u << lane) - 1u));
                        if (in) {
                            if (i < lo + m) {
                                if (!f) scr[lo + (i - lo) - pre] = i;
                            } else if (f) {
                                scr[hi - m + pre] = i;
                            }
                        }
                        base += __popc(mask);
                    }
                    // g = number of misplaced elements in the left part = m - #('<' elements inside the left part)
                    u32 inl = 0;
                    for (u32 i = lo + lane; i < lo + m; i += 32) inl += (col[i] < val) ? 1u : 0u;
                    const u32 gcount = m - __reduce_add_sync(FULL, inl);
                    __syncwarp();
                    for (u32 kk = lane; kk < gcount; kk += 32) {
                        const u32 pa = scr[lo + kk], pb = scr[hi - 1 - kk];
#pragma unroll
                        for (int c = 0; c < DIM; ++c)
                            if (c < (int)dim) {
                                float *cc = q + (size_t)c * npad;
                                const float xa = cc[pa], xb = cc[pb];
                                cc[pa] = xb;
                                cc[pb] = xa;
                            }
                        const u32 ia = perm[pa], ib = perm[pb];
                        perm[pa] = ib;
                        perm[pb] = ia;
                    }
                    __syncwarp();
                }
                if (lane == 0) {
                    if (lim >= 2) {
                        const u32 o = atomicAdd(&s_cnt[cur ^ 1], 1u);
                        list[cur ^ 1][2 * o] = lo;
                        list[cur ^ 1][2 * o + 1] = lo + lim;
                    }
                    if (count - lim >= 2) {
                        const u32 o = atomicAdd(&s_cnt[cur ^ 1], 1u);
                        list[cur ^ 1][2 * o] = lo + lim;
                        list[cur ^ 1][2 * o + 1] = hi;
                    }
                }
            }
            __threadfence_block();
            __syncthreads();
            if (tid == 0) s_cnt[cur] = 0;
            cur ^= 1;
            __syncthreads();
        }
        // the permuted cloud, row-major
        float *rows = a.rows + (size_t)cloud * n * dim;
        for (u32 f = tid; f < n * dim; f += KT_T) {
            const u32 i = f / dim, c = f - i * dim;
            rows[f] = q[(size_t)c * npad + i];
        }
        __syncthreads();
    }
}

// positions (what the vanilla kernels returned over the permuted rows) -> original ids (wrapper.hpp:39-41)
__global__ void kdtree_map_kernel(u64 *out, const unsigned char *region, size_t region_stride, u32 B, u32 k, u32 dim, u32 npad) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * k) return;
    const u32 *perm = reinterpret_cast<const u32 *>(region + (i / k) * region_stride) + (size_t)(dim + 1) * npad;
    out[i] = (u64)perm[(u32)out[i]];
}

size_t kdtree_region_bytes(size_t n, size_t dim) {
    const size_t npad = (n + 31) & ~(size_t)31;
    return (((dim + 4) * npad) * 4 + 255) & ~(size_t)255;
}

static int pad_dim_t(int dim) { return dim <= 2 ? 2 : dim == 3 ? 3 : dim == 4 ? 4 : dim <= 6 ? 6 : 8; }

cudaError_t launch_kdtree_build(const float *pts, unsigned char *region, size_t region_stride, float *rows, u32 B, u32 n,
                                u32 dim, int n_sms, cudaStream_t st) {
    KdtreeArgs a;
    a.pts = pts;
    a.region = region;
    a.region_stride = region_stride;
    a.rows = rows;
    a.B = B;
    a.n = n;
    a.npad = (n + 31) & ~31u;
    a.dim = dim;
    const u32 grid = B < (u32)n_sms ? B : (u32)n_sms;
    switch (pad_dim_t((int)dim)) {
        case 2: kdtree_build_kernel<2><<<grid, KT_T, 0, st>>>(a); break;
        case 3: kdtree_build_kernel<3><<<grid, KT_T, 0, st>>>(a); break;
        case 4: kdtree_build_kernel<4><<<grid, KT_T, 0, st>>>(a); break;
        case 6: kdtree_build_kernel<6><<<grid, KT_T, 0, st>>>(a); break;
        default: kdtree_build_kernel<8><<<grid, KT_T, 0, st>>>(a); break;
    }
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_kdtree_map(u64 *out, const unsigned char *region, size_t region_stride, u32 B, u32 n, u32 k, u32 dim,
                              cudaStream_t st) {
    const size_t tot = (size_t)B * k;
    kdtree_map_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(out, region, region_stride, B, k, dim, (n + 31) & ~31u);
    count_launch();
    return cudaGetLastError();
}

}  // namespace fps
