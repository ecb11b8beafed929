// seqsum.cu -- test / inspection entry for the tile-parallel sequential binary32 sum (seqsum.cuh): one warp sums a
// device array exactly as `float s = 0; for (x : a) s += x;` would (reference src/_ext/KDTreeBase.h:151-158), through
// the same tile code the kd build kernels use.  include/fps_b200.h: fps_b200_seqsum_dev.
#include "common.cuh"
#include "engine.h"
#include "seqsum.cuh"

namespace fps {

constexpr u32 SQ_CH = 8192;   // floats staged per chunk
template <int EPL>
__global__ void __launch_bounds__(32) seqsum_kernel(const float *x, size_t n, float *out, u32 *fast_tiles) {
    __shared__ __align__(16) float buf[SQ_CH];
    float sum = 0.0f;
    u32 fast = 0, hint = 0;
    for (size_t i0 = 0; i0 < n; i0 += SQ_CH) {
        const u32 m = (u32)((n - i0 < SQ_CH) ? n - i0 : SQ_CH);
        for (u32 i = threadIdx.x; i < m; i += 32) buf[i] = x[i0 + i];
        __syncwarp();
        const u32 a = smem_u32(buf);
        constexpr u32 TILE = 32 * EPL;
        u32 i = 0;
        for (; i + TILE <= m; i += TILE) {
            if (seq_sum_tile<EPL>(a + 4 * i, sum, hint)) ++fast;
            else sum = sq_chain16(a + 4 * i, TILE, sum);
        }
        for (; i < m; ++i) sum = __fadd_rn(sum, buf[i]);
        __syncwarp();
    }
    if (threadIdx.x == 0) {
        *out = sum;
        if (fast_tiles) *fast_tiles = fast;
    }
}

// the two-phase form (kdbuild.cu: gb_psum_a / gb_psum_b / gb_split) on one column, one CTA: every warp prepares tiles
// (double-precision tile sums, then the integer record under the binade guessed from their prefix), warp 0 walks the records
constexpr u32 SQ_MAXT = 1u << 14;   // tiles (8 M values)
__device__ double g_sq_dsum[SQ_MAXT];
__device__ SeqTileRec g_sq_recs[SQ_MAXT];
__global__ void __launch_bounds__(256) seqsum_records_kernel(const float *x, size_t n, float *out, u32 *fast_tiles) {
    __shared__ __align__(16) float tile[8][512];
    const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const u32 ntile = (u32)(n / 512);
    for (u32 t = warp; t < ntile; t += 8) {
        double acc = 0.0;
        for (u32 k = 0; k < 16; ++k) acc += (double)x[(size_t)t * 512 + k * 32 + lane];
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
        if (lane == 0) g_sq_dsum[t] = acc;
    }
    __syncthreads();
    for (u32 t = warp; t < ntile; t += 8) {
        for (u32 k = 0; k < 16; ++k) tile[warp][k * 32 + lane] = x[(size_t)t * 512 + k * 32 + lane];
        double pre = 0.0;
        for (u32 u = lane; u < t; u += 32) pre += g_sq_dsum[u];
        for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(FULL, pre, o);
        const u32 ef = (__float_as_uint((float)pre) >> 23) & 0xffu;
        __syncwarp();
        int t0 = 0, t1 = 0, lo = 0, hi = 0;
        const bool ok = seq_sum_tile_record<16>(smem_u32(tile[warp]), ef, t0, t1, lo, hi);
        if (lane == 0) {
            SeqTileRec r;
            r.ef = ok ? ef : 0u, r.tot0 = t0, r.tot1 = t1, r.lo = lo, r.hi = hi, r.pad[0] = r.pad[1] = r.pad[2] = 0u;
            g_sq_recs[t] = r;
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp) return;
    float sum = 0.0f;
    u32 fast = 0;
    for (u32 t = 0; t < ntile; ++t) {
        const SeqTileRec r = g_sq_recs[t];
        if (seq_sum_apply_record(sum, r.ef, r.tot0, r.tot1, r.lo, r.hi)) {
            ++fast;
            continue;
        }
        for (u32 k = 0; k < 16; ++k) tile[0][k * 32 + lane] = x[(size_t)t * 512 + k * 32 + lane];
        __syncwarp();
        sum = sq_chain16(smem_u32(tile[0]), 512, sum);
        __syncwarp();
    }
    for (size_t i = (size_t)ntile * 512; i < n; ++i) sum = __fadd_rn(sum, x[i]);
    if (lane == 0) {
        *out = sum;
        if (fast_tiles) *fast_tiles = fast;
    }
}

cudaError_t launch_seqsum(const float *x, size_t n, float *out, u32 *fast_tiles, int epl, cudaStream_t st) {
    if (epl < 0) {   // two-phase
        if (n / 512 > SQ_MAXT) return cudaErrorNotSupported;
        seqsum_records_kernel<<<1, 256, 0, st>>>(x, n, out, fast_tiles);
        count_launch();
        return cudaGetLastError();
    }
    if (epl == 8) seqsum_kernel<8><<<1, 32, 0, st>>>(x, n, out, fast_tiles);
    else seqsum_kernel<16><<<1, 32, 0, st>>>(x, n, out, fast_tiles);
    count_launch();
    return cudaGetLastError();
}

}  // namespace fps
