// microbenchmark: latencies of the primitives the per-pick critical path is built from (B200, sm_100a):
// redux.sync, shfl, ballot, LDS, bar.sync, 64-bit warp_max_key, and TMEM (tcgen05.ld/st) used as a
// lane-private dynamically indexed scratchpad.  One warp unless stated.  Prints cycles per dependent op.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../fpsample_b200/csrc/common.cuh"
using namespace fps;

#define N_IT 256

__global__ void k_redux(u32 *out, long long *cyc, u32 seed) {
    u32 x = seed + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N_IT; ++i) x = __reduce_max_sync(FULL, x) + (threadIdx.x & 1);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_reduxf(float *out, long long *cyc, float seed) {
    float x = seed + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N_IT; ++i) {
        float m;
        asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(x));
        x = m + (float)(threadIdx.x & 1);
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
}
__global__ void k_shfl(u32 *out, long long *cyc, u32 seed) {
    u32 x = seed + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N_IT; ++i) x = __shfl_sync(FULL, x, (x + 1) & 31) + 1;
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
}
__global__ void k_ballot(u32 *out, long long *cyc, u32 seed) {
    u32 x = seed + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N_IT; ++i) x = __ffs(__ballot_sync(FULL, (x & 7) == 3)) + x;
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[3] = t1 - t0;
}
__global__ void k_lds(u32 *out, long long *cyc, u32 seed) {
    __shared__ u32 tab[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) tab[i] = (i * 33 + 7 + seed) & 1023;
    __syncwarp();
    u32 x = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N_IT; ++i) x = tab[x];
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[4] = t1 - t0;
}
__global__ void k_bar(u32 *out, long long *cyc, int slot) {
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N_IT; ++i) __syncthreads();
    long long t1 = clock64();
    out[threadIdx.x] = 0;
    if (threadIdx.x == 0) cyc[slot] = t1 - t0;
}
__global__ void k_maxkey(u64 *out, long long *cyc, u64 seed) {
    u64 x = seed * (threadIdx.x + 1);
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N_IT; ++i) x = warp_max_key(x) + threadIdx.x;
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[9] = t1 - t0;
}
// bar.sync + smem exchange round trip: write my value, barrier, read neighbour warp's value (what a per-pick
// cross-warp arg-max costs), 128 threads
__global__ void k_xchg(u32 *out, long long *cyc) {
    __shared__ u32 slot[2][32];
    u32 x = threadIdx.x;
    const u32 w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N_IT; ++i) {
        if ((threadIdx.x & 31) == 0) slot[i & 1][w] = x;
        __syncthreads();
        x = slot[i & 1][(w + 1) % nw] + 1;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[10] = t1 - t0;
}

// ---- TMEM as scratch ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tm_ld4(u32 taddr, u32 &a, u32 &b, u32 &c, u32 &d) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_ld1(u32 taddr, u32 &a) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(a) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_st4(u32 taddr, u32 a, u32 b, u32 c, u32 d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tm_st1(u32 taddr, u32 a) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(a) : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 4 warps: each fills its lane quarter (512 columns) with a pattern, reads it back and checks; then a TMEM
// pointer chase (latency) and a burst of independent x4 loads (throughput).
__global__ void __launch_bounds__(128, 1) k_tmem(u32 *out, long long *cyc, u32 *errs) {
    __shared__ u32 tbase;
    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const u32 base = tbase + ((warp * 32u) << 16);   // lane field = bits 31:16
    // fill: column c of lane l = (warp<<28) | (l << 16) | c'  (c' = a chase target for c % 4 == 0)
    for (u32 c = 0; c < 512; c += 4) {
        const u32 nxt = ((c / 4) * 37 + 11) % 128 * 4;
        tm_st4(base + c, nxt, (warp << 28) | (lane << 16) | (c + 1), (warp << 28) | (lane << 16) | (c + 2), (warp << 28) | (lane << 16) | (c + 3));
    }
    tm_wait_st();
    u32 bad = 0;
    for (u32 c = 0; c < 512; c += 4) {
        u32 a, b, d, e;
        tm_ld4(base + c, a, b, d, e);
        tm_wait_ld();
        const u32 nxt = ((c / 4) * 37 + 11) % 128 * 4;
        bad += (a != nxt) + (b != ((warp << 28) | (lane << 16) | (c + 1))) + (d != ((warp << 28) | (lane << 16) | (c + 2))) +
               (e != ((warp << 28) | (lane << 16) | (c + 3)));
    }
    // single-column overwrite then read back through x4 (what the running-distance update does)
    tm_st1(base + 7, 0xabcd0000u | lane);
    tm_wait_st();
    {
        u32 a, b, d, e;
        tm_ld4(base + 4, a, b, d, e);
        tm_wait_ld();
        bad += (e != (0xabcd0000u | lane));
    }
    atomicAdd(errs, bad);
    // latency: pointer chase
    u32 x = 0;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N_IT; ++i) {
        u32 a, b, d, e;
        tm_ld4(base + x, a, b, d, e);
        tm_wait_ld();
        x = a;
    }
    long long t1 = clock64();
    // throughput: 16 independent x4 loads per wait
    u32 acc = 0;
    long long t2 = clock64();
#pragma unroll 1
    for (int i = 0; i < N_IT / 16; ++i) {
        u32 v[16][4];
#pragma unroll
        for (int j = 0; j < 16; ++j) tm_ld4(base + ((i * 16 + j) * 4 & 511), v[j][0], v[j][1], v[j][2], v[j][3]);
        tm_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) acc += v[j][0] ^ v[j][1] ^ v[j][2] ^ v[j][3];
    }
    long long t3 = clock64();
    // store -> wait -> load round trip (running-distance update then reread)
    long long t4 = clock64();
#pragma unroll 4
    for (int i = 0; i < N_IT; ++i) {
        tm_st1(base + 9, x);
        tm_wait_st();
        u32 a;
        tm_ld1(base + 9, a);
        tm_wait_ld();
        x = a + 1;
    }
    long long t5 = clock64();
    out[threadIdx.x] = x + acc;
    if (threadIdx.x == 0) {
        cyc[11] = t1 - t0;
        cyc[12] = t3 - t2;
        cyc[13] = t5 - t4;
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

int main() {
    u32 *out, *errs;
    long long *cyc;
    cudaMalloc(&out, 1 << 16);
    cudaMallocManaged(&cyc, 256);
    cudaMallocManaged(&errs, 4);
    *errs = 0;
    for (int rep = 0; rep < 2; ++rep) {
        k_redux<<<1, 32>>>(out, cyc, 5);
        k_reduxf<<<1, 32>>>((float *)out, cyc, 5.f);
        k_shfl<<<1, 32>>>(out, cyc, 5);
        k_ballot<<<1, 32>>>(out, cyc, 5);
        k_lds<<<1, 32>>>(out, cyc, 5);
        k_bar<<<1, 64>>>(out, cyc, 5);
        k_bar<<<1, 128>>>(out, cyc, 6);
        k_bar<<<1, 256>>>(out, cyc, 7);
        k_bar<<<1, 512>>>(out, cyc, 8);
        k_maxkey<<<1, 32>>>((u64 *)out, cyc, 12345);
        k_xchg<<<1, 128>>>(out, cyc);
        *errs = 0;
        k_tmem<<<1, 128>>>(out, cyc, errs);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("CUDA error: %s\n", cudaGetErrorString(e));
            return 1;
        }
    }
    const double n = N_IT;
    printf("redux.max.u32 %.1f | redux.max.f32 %.1f | shfl %.1f | ballot+ffs %.1f | LDS chase %.1f cyc per dependent op\n", cyc[0] / n,
           cyc[1] / n, cyc[2] / n, cyc[3] / n, cyc[4] / n);
    printf("bar.sync 64/128/256/512 threads: %.1f %.1f %.1f %.1f | warp_max_key(u64) %.1f | STS+bar+LDS exchange (128 thr) %.1f\n",
           cyc[5] / n, cyc[6] / n, cyc[7] / n, cyc[8] / n, cyc[9] / n, cyc[10] / n);
    printf("TMEM: readback errors %u | ld.x4 chase %.1f cyc | ld.x4 burst %.1f cyc per load (4 warps active) | st1+wait+ld1+wait %.1f cyc\n",
           *errs, cyc[11] / n, cyc[12] / n, cyc[13] / n);
    return 0;
}
