// microbenchmark: cycles per dependent FADD (register chain) and for the staged sequential-sum loop
#include <cstdio>
#include <cuda_runtime.h>
#include "../../fpsample_b200/csrc/kdcommon.cuh"
using namespace fps;
__global__ void k_reg(float *out, long long *cyc, float x) {
    float s = 0.f;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; ++i) {
#pragma unroll
        for (int j = 0; j < 32; ++j) s = __fadd_rn(s, x);
    }
    long long t1 = clock64();
    out[threadIdx.x] = s; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_staged(const float *src, u32 n, float *out, long long *cyc) {
    __shared__ __align__(16) float buf[256];
    long long t0 = clock64();
    float s = seq_sum_staged(src, n, buf);
    long long t1 = clock64();
    out[threadIdx.x] = s; if (threadIdx.x == 0) cyc[1] = t1 - t0;
}
__global__ void k_shfl(const float *src, u32 n, float *out, long long *cyc) {
    long long t0 = clock64();
    float s = seq_sum(src, n);
    long long t1 = clock64();
    out[threadIdx.x] = s; if (threadIdx.x == 0) cyc[2] = t1 - t0;
}
__global__ void k_tma(const float *src, u32 n, float *out, long long *cyc) {
    __shared__ __align__(16) float ring[SS_STAGES * SS_TILE];
    __shared__ u64 bars[SS_STAGES];
    if (threadIdx.x < SS_STAGES) mbar_init(smem_u32(&bars[threadIdx.x]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    u32 phase = 0;
    long long t0 = clock64();
    float s = seq_sum_tma(src, n, ring, bars, phase);
    long long t1 = clock64();
    out[threadIdx.x] = s; if (threadIdx.x == 0) cyc[3] = t1 - t0;
}
__global__ void k_smem(float *out, long long *cyc) {
    __shared__ __align__(16) float buf[8192];
    for (int i = threadIdx.x; i < 8192; i += 32) buf[i] = 0.f;
    __syncwarp();
    long long t0 = clock64();
    float s = seq_sum_smem(buf, 8192);
    long long t1 = clock64();
    out[threadIdx.x] = s; if (threadIdx.x == 0) cyc[4] = t1 - t0;
}
int main() {
    const u32 n = 1 << 20;
    float *src, *out; long long *cyc;
    cudaMalloc(&src, n * 4); cudaMalloc(&out, 4096); cudaMallocManaged(&cyc, 64);
    cudaMemset(src, 0, n * 4);
    for (int rep = 0; rep < 2; ++rep) {
        k_reg<<<1, 32>>>(out, cyc, 1.0f); k_staged<<<1, 32>>>(src, n, out, cyc); k_shfl<<<1, 32>>>(src, n, out, cyc); k_tma<<<1, 32>>>(src, n, out, cyc); k_smem<<<1, 32>>>(out, cyc);
        cudaDeviceSynchronize();
    }
    printf("reg chain: %.2f cyc/add   staged: %.2f cyc/elem   shfl: %.2f cyc/elem   tma ring: %.2f cyc/elem   smem direct: %.2f cyc/elem\n", cyc[0] / 32768.0, cyc[1] / (double)n, cyc[2] / (double)n, cyc[3] / (double)n, cyc[4] / 8192.0);
    return 0;
}
