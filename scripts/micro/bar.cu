// microbenchmark: a short phase executed by a few warps of a 1024-thread CTA between two __syncthreads (B200):
// what does the barrier cost when 27 warps arrive early?  Variants: with / without a global-memory spin before it.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../fpsample_b200/csrc/common.cuh"
using namespace fps;
#define N_IT 200
#define R4(x) x x x x
#define R16(x) R4(R4(x))
#define R128(x) R16(x) R16(x) R16(x) R16(x) R16(x) R16(x) R16(x) R16(x)
#define OP "xor.b32 %0, %0, %1; add.u32 %0, %0, 3;\n"
#ifndef FILL_KB
#define FILL_KB 0
#endif
__global__ void __launch_bounds__(1024, 1) k(long long *cyc, u32 *out, u64 *gk_init, int mode, u32 *gmem, u32 gsz) {
    extern __shared__ u64 sm[];
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (u32 i = tid; i < 4096; i += 1024) sm[i] = gk_init[i];
    __syncthreads();
    long long acc[3] = {0, 0, 0};
    u64 keep = 0;
    for (int it = 0; it < N_IT; ++it) {
        if (mode >= 1 && tid < 720) {   // a dependent global round trip first (like the gather)
            u32 v;
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(gmem + (tid + it) % 1024) : "memory");
            sm[2048 + tid] += v;
        }
        if (mode >= 0) {   // FILL_KB of straight-line code between the phases (instruction footprint of a big loop body)
            u32 x = (u32)keep + tid;
#pragma unroll
            for (int b = 0; b < FILL_KB / 4; ++b) asm volatile(R128(OP) : "+r"(x) : "r"(tid));
            keep += x;
        }
        __syncthreads();
        const long long t0 = clock64();
        if (warp < 5) {
            u64 b8 = 0, b4 = 0, b1 = 0;
            if (tid < gsz) {
                const u64 *kc = sm + tid * 10;
                const u64 wb = kc[9];
                b8 = wb > kc[8] ? wb : kc[8];
                b4 = wb > kc[4] ? wb : kc[4];
                b1 = wb > kc[1] ? wb : kc[1];
            }
            b8 = warp_max_key(b8); b4 = warp_max_key(b4); b1 = warp_max_key(b1);
            if (lane == 0) { sm[3000 + warp * 4] = b8; sm[3001 + warp * 4] = b4; sm[3002 + warp * 4] = b1; }
        }
        __syncthreads();
        const long long t1 = clock64();
        u64 B = 0;
        for (u32 w = 0; w < 5; ++w) B = B > sm[3000 + w * 4] ? B : sm[3000 + w * 4];
        keep += B;
        if (mode == 2) __syncthreads();
        const long long t2 = clock64();
        acc[0] += t1 - t0; acc[1] += t2 - t1;
    }
    out[blockIdx.x * 1024 + tid] = (u32)keep;
    if (tid == 0 && blockIdx.x == 0) { cyc[0] = acc[0] / N_IT; cyc[1] = acc[1] / N_IT; }
}
int main() {
    long long *cyc; u32 *out, *gmem; u64 *init;
    cudaMallocManaged(&cyc, 64); cudaMalloc(&out, 148 * 4096); cudaMalloc(&gmem, 4096); cudaMallocManaged(&init, 4096 * 8);
    for (int i = 0; i < 4096; ++i) init[i] = (u64)i * 0x9e3779b97f4a7c15ull;
    cudaMemset(gmem, 0, 4096);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    for (int grid : {1, 144})
        for (int mode = 0; mode < 3; ++mode) {
            k<<<grid, 1024, 160 * 1024>>>(cyc, out, init, mode, gmem, 144);
            cudaError_t e = cudaDeviceSynchronize();
            printf("fill %d KB grid %3d mode %d: bounds phase %lld cycles, reduce %lld (%s)\n", FILL_KB, grid, mode, cyc[0], cyc[1], cudaGetErrorString(e));
        }
    return 0;
}
