// microbenchmark: the grid-wide exchange of kdline_grid_kernel in isolation (B200): every CTA publishes CH stamped
// 16-byte chunks per round, every CTA gathers all G*CH chunks.  Prints cycles per round for variants of the gather.
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include "../../fpsample_b200/csrc/common.cuh"
using namespace fps;
#define CH 19
#define ROUNDS 400
__device__ __forceinline__ uint4 ldr(const uint4 *p) {
    uint4 v; asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void str(uint4 *p, u32 x, u32 y, u32 z, u32 w) {
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
// mode 0: thread c polls header of CTA c then fetches its chunks 4 at a time; mode 1: one thread per chunk polls (G*CH threads, strided)
// mode 2: like 0 but the pollers back off with nanosleep(64)
__global__ void __launch_bounds__(1024, 1) k(uint4 *pub, long long *cyc, int mode, u32 *sink) {
    extern __shared__ uint4 gbuf[];
    const u32 tid = threadIdx.x, G = gridDim.x, cta = blockIdx.x;
    u32 acc = 0;
    long long tb = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (u32 round = 0; round < ROUNDS; ++round) {
        const u32 stamp = round + 1;
        uint4 *dst = pub + ((size_t)(round & 1) * G + cta) * CH;
        if (tid >= 32 && tid < 32 + CH) str(dst + (tid - 32), stamp, tid, cta, round);
        if (mode == 3 || mode == 4) {   // warp-uniform spin: the whole warp leaves the loop together
            for (u32 i0 = 0; i0 < G * CH; i0 += 1024) {
                const u32 i = i0 + tid;
                const bool act = i < G * CH;
                const uint4 *src = pub + (size_t)(round & 1) * G * CH + (act ? i : 0);
                uint4 v;
                bool ok;
                do { v = ldr(src); ok = !act || v.x == stamp; } while (!__all_sync(0xffffffffu, ok));
                if (act) gbuf[i] = v;
            }
            if (mode == 4) __syncwarp();
        } else if (mode == 1) {
            for (u32 i = tid; i < G * CH; i += 1024) {
                const uint4 *src = pub + (size_t)(round & 1) * G * CH + i;
                uint4 v; do { v = ldr(src); } while (v.x != stamp);
                gbuf[i] = v;
            }
        } else if (tid < G) {
            const uint4 *src = pub + ((size_t)(round & 1) * G + tid) * CH;
            uint4 v; do { v = ldr(src); if (mode == 2 && v.x != stamp) __nanosleep(64); } while (v.x != stamp);
            gbuf[tid * CH] = v;
            for (u32 ch = 1; ch < CH; ch += 4) {
                uint4 w[4];
#pragma unroll
                for (u32 x = 0; x < 4; ++x) if (ch + x < CH) w[x] = ldr(src + ch + x);
#pragma unroll
                for (u32 x = 0; x < 4; ++x) if (ch + x < CH) { while (w[x].x != stamp) w[x] = ldr(src + ch + x); gbuf[tid * CH + ch + x] = w[x]; }
            }
        }
        __syncthreads();
        const long long b0 = clock64();
        if (tid < 160) {   // the "bounds" phase of kdline_grid_kernel: 5 warps, a few shared loads, six redux
            u64 b8 = 0, b4 = 0;
            if (tid < G) { const uint4 x = gbuf[tid * CH], y = gbuf[tid * CH + 1]; b8 = ((u64)x.y << 32) | x.z; b4 = ((u64)y.y << 32) | y.z; }
            b8 = warp_max_key(b8); b4 = warp_max_key(b4); b8 = warp_max_key(b8 + b4);
            if ((tid & 31) == 0) ((u64 *)(gbuf + G * CH))[tid >> 5] = b8;
        }
        __syncthreads();
        const long long b1 = clock64();
        tb += b1 - b0;
        acc += gbuf[(round * 7 + tid) % (G * CH)].y + (u32)((u64 *)(gbuf + G * CH))[tid & 3];
        __syncthreads();
    }
    const long long t1 = clock64();
    if (tid == 0) { cyc[cta] = (t1 - t0) / ROUNDS; cyc[160 + cta] = tb / ROUNDS; }
    sink[cta * 1024 + tid] = acc;
}
int main() {
    const int G = 144;
    uint4 *pub; long long *cyc; u32 *sink;
    cudaMalloc(&pub, 2 * G * CH * 16); cudaMallocManaged(&cyc, 400 * 8); cudaMalloc(&sink, G * 1024 * 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 190 * 1024);
    for (int mode = 0; mode < 5; ++mode) {
        cudaMemset(pub, 0, 2 * G * CH * 16);
        void *args[] = {&pub, &cyc, &mode, &sink};
        cudaLaunchCooperativeKernel((void *)k, dim3(G), dim3(1024), args, 190 * 1024, 0);
        cudaError_t e = cudaDeviceSynchronize();
        long long mx = 0, mn = 1ll << 60; for (int i = 0; i < G; ++i) { mx = cyc[i] > mx ? cyc[i] : mx; mn = cyc[i] < mn ? cyc[i] : mn; }
        printf("mode %d: cycles per round min %lld max %lld, bounds phase after the gather (CTA 0) %lld (%s)\n", mode, mn, mx, cyc[160], cudaGetErrorString(e));
    }
    return 0;
}
