// warpops.cu -- dependent-issue latency (cycles per op in a dependent chain, one warp) of the warp-collective and ALU
// ops on the one-warp-per-cloud sampler's pick path.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warpops warpops.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned int u32;
#define FULL 0xffffffffu
template <int OP>
__global__ void k(u32 *out, long long *cyc, int iters, u32 seed) {
    __shared__ u32 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) sm[i] = (i * 33 + 7) & 1023;
    __syncwarp();
    u32 x = seed + threadIdx.x;
    float f = (float)x;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (OP == 0) x = __reduce_max_sync(FULL, x) + threadIdx.x;
        if (OP == 1) x = __shfl_sync(FULL, x, (x + 1) & 31) + 1;
        if (OP == 2) x = __ballot_sync(FULL, x & 1) + threadIdx.x;
        if (OP == 3) x = sm[x & 1023];
        if (OP == 4) { f = fminf(f, __uint_as_float(x)) + 1.0f; x += 1; }
        if (OP == 5) x = __reduce_min_sync(FULL, __reduce_max_sync(FULL, x) == x ? threadIdx.x : 99u) + x;
        if (OP == 6) { f = __fadd_rn(f, 1.0f); }
        if (OP == 7) { x = max(x, seed) + 1; }
        if (OP == 8) { x = __popc(x) + x; }
        if (OP == 9) { x = (x & 1) ? __shfl_sync(FULL, x, 3) : x + 1; }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + (u32)f;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    u32 *o; long long *c; cudaMalloc(&o, 128); cudaMalloc(&c, 8);
    const char *names[] = {"redux.max", "shfl (dependent lane)", "ballot", "lds (pointer chase)", "fmnmx+fadd", "redux.max + redux.min (arg-max pair)", "fadd", "imax+iadd", "popc+iadd", "branchy shfl"};
    const int iters = 4000;
    long long h;
#define RUN(OP) k<OP><<<1, 32>>>(o, c, iters, 5); cudaDeviceSynchronize(); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); printf("%-40s %.1f cycles per iteration\n", names[OP], (double)h / iters);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9)
    return 0;
}
