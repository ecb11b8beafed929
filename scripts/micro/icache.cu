// microbenchmark: cost of executing a loop body that does not fit the instruction caches (B200): a dependent chain
// of N integer ops (16 bytes of SASS each), body sizes 2..64 KB, one warp and eight warps (two per scheduler).
#include <cstdio>
#include <cuda_runtime.h>
#define R4(x) x x x x
#define R16(x) R4(R4(x))
#define R128(x) R16(x) R16(x) R16(x) R16(x) R16(x) R16(x) R16(x) R16(x)
#define OP "xor.b32 %0, %0, %1; add.u32 %0, %0, 3;\n"
template <int KB>
__global__ void k(unsigned *out, long long *cyc, unsigned seed, int iters) {
    unsigned x = seed + threadIdx.x;
    long long t0 = 0;
    for (int it = 0; it < iters + 2; ++it) {
        if (it == 2) t0 = clock64();
#pragma unroll
        for (int b = 0; b < KB / 4; ++b) asm volatile(R128(OP) : "+r"(x) : "r"(seed));   // 128 x 2 instrs x 16 B = 4 KB
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = (t1 - t0) / iters;
}
template <int KB>
void run(unsigned *out, long long *cyc) {
    for (int warps : {1, 8, 32}) {
        k<KB><<<1, 32 * warps>>>(out, cyc, 5, 50);
        cudaDeviceSynchronize();
        printf("body %2d KB, %2d warps: %6lld cycles per iteration = %.2f cycles per instruction\n", KB, warps, cyc[0], (double)cyc[0] / (KB * 64));
    }
}
int main() {
    unsigned *out; long long *cyc;
    cudaMalloc(&out, 4096 * 4); cudaMallocManaged(&cyc, 8);
    run<4>(out, cyc); run<8>(out, cyc); run<16>(out, cyc); run<24>(out, cyc); run<32>(out, cyc); run<48>(out, cyc); run<64>(out, cyc);
    return 0;
}
