// microbenchmark: what the selector warps of kdline_grid_kernel pay per primitive while the other warps of a
// 1024-thread CTA wait at barrier 0 (B200, sm_100a).  Prints cycles per op for thread 0.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../fpsample_b200/csrc/common.cuh"
using namespace fps;
#define N_IT 64
__device__ __forceinline__ void bar_named(u32 id, u32 n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__global__ void __launch_bounds__(1024, 1) k(long long *cyc, u32 *out, u32 seed, int others_wait, u32 *gflag) {
    extern __shared__ u32 sm[];
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 64) sm[tid] = 0;
    __syncthreads();
    if (warp >= 8) {
        if (others_wait == 1) __syncthreads();
        if (others_wait == 2) { while (((volatile u32 *)gflag)[0] == 0) {} }   // spin on global memory
        return;
    }
    long long t[8];
    u32 x = seed + tid;
    t[0] = clock64();
    for (int i = 0; i < N_IT; ++i) bar_named(1, 256);
    t[1] = clock64();
    u64 kx = ((u64)x << 32) | tid;
    for (int i = 0; i < N_IT; ++i) kx = warp_max_key(kx) + lane;
    t[2] = clock64();
    for (int i = 0; i < N_IT; ++i) {
        u32 b = 0;
        if (lane == 0) b = atomicAdd(&sm[0], x & 7);
        x += __shfl_sync(FULL, b, 0);
    }
    t[3] = clock64();
    for (int i = 0; i < N_IT; ++i) x = sm[32 + (x & 31)] + x;
    t[4] = clock64();
    for (int i = 0; i < N_IT; ++i) x += __popc(__ballot_sync(FULL, (x >> (i & 7)) & 1));
    t[5] = clock64();
    for (int i = 0; i < N_IT; ++i) x = __reduce_max_sync(FULL, x) + lane;
    t[6] = clock64();
    out[tid] = x + (u32)kx;
    if (tid == 0) { for (int i = 0; i < 6; ++i) cyc[i] = (t[i + 1] - t[i]) / N_IT; gflag[0] = 1; }
    if (others_wait == 1) __syncthreads();
}
int main() {
    long long *cyc; u32 *out, *gf;
    cudaMallocManaged(&cyc, 64); cudaMalloc(&out, 4096); cudaMalloc(&gf, 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 190 * 1024);
    const char *names[] = {"others exit", "others at __syncthreads", "others spin on global"};
    for (int m = 0; m < 3; ++m) {
        cudaMemset(gf, 0, 4);
        k<<<1, 1024, 190 * 1024>>>(cyc, out, 7, m, gf);
        cudaError_t e = cudaDeviceSynchronize();
        printf("%-26s: named bar(256) %lld | warp_max_key %lld | ATOMS+shfl (8 warps, 1 addr) %lld | LDS dep %lld | ballot+popc %lld | redux %lld  (%s)\n",
               names[m], cyc[0], cyc[1], cyc[2], cyc[3], cyc[4], cyc[5], cudaGetErrorString(e));
    }
    return 0;
}
