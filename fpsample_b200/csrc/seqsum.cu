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

cudaError_t launch_seqsum(const float *x, size_t n, float *out, u32 *fast_tiles, int epl, cudaStream_t st) {
    if (epl == 8) seqsum_kernel<8><<<1, 32, 0, st>>>(x, n, out, fast_tiles);
    else seqsum_kernel<16><<<1, 32, 0, st>>>(x, n, out, fast_tiles);
    count_launch();
    return cudaGetLastError();
}

}  // namespace fps
