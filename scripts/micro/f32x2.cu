// f32x2.cu -- does Blackwell's packed binary32 arithmetic (add/sub/mul.rn.f32x2 -> FADD2/FMUL2) help the bucket
// pass of the one-warp-per-cloud sampler?  Same loop as kdline_warp.cu's pending-sample loop: 8 chunks x 3
// coordinates in registers, nref samples from shared memory, v = min(v, |x - ref|^2) with individually rounded
// ops.  Reports cycles per (ref x 8 chunks) for W warps on one SM, scalar vs packed, and checks bit equality.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o f32x2 f32x2.cu
#include <cstdio>
#include <cstdint>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up(u64 v, float &a, float &b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
// ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under --fmad=false, so the product is written as
// fma(a, b, -0.0): one rounding of a*b, and adding -0 changes nothing (+0 + -0 = +0, -0 + -0 = -0); an FMA result
// cannot be contracted into the following add
__device__ __forceinline__ u64 mul2(u64 a, u64 b, u64 nz) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(nz)); return r; }

template <int MODE>
__global__ void k(float *out, const float *in, int nref, int reps, long long *cyc, u64 nz) {
    __shared__ float4 refs[64];
    if (threadIdx.x < 64) refs[threadIdx.x] = make_float4(in[threadIdx.x * 3], in[threadIdx.x * 3 + 1], in[threadIdx.x * 3 + 2], 0.f);
    __syncthreads();
    float x[3][8], v[8];
    for (int c = 0; c < 3; ++c)
        for (int u = 0; u < 8; ++u) x[c][u] = in[200 + (c * 8 + u) * 32 + (threadIdx.x & 31)];
    for (int u = 0; u < 8; ++u) v[u] = 3.0e38f;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (MODE == 0) {
            for (int i = 0; i < nref; ++i) {
                const float4 f = refs[(i + r) & 63];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    float d0 = __fsub_rn(x[0][u], f.x), d1 = __fsub_rn(x[1][u], f.y), d2 = __fsub_rn(x[2][u], f.z);
                    float s = __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
                    v[u] = fminf(v[u], s);
                }
            }
        } else {
            u64 X[3][4];
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int u = 0; u < 4; ++u) X[c][u] = pk(x[c][2 * u], x[c][2 * u + 1]);
            for (int i = 0; i < nref; ++i) {
                const float4 f = refs[(i + r) & 63];
                const u64 fx = pk(f.x, f.x), fy = pk(f.y, f.y), fz = pk(f.z, f.z);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    u64 d0 = sub2(X[0][u], fx), d1 = sub2(X[1][u], fy), d2 = sub2(X[2][u], fz);
                    u64 s = add2(add2(mul2(d0, d0, nz), mul2(d1, d1, nz)), mul2(d2, d2, nz));
                    float sa, sb;
                    up(s, sa, sb);
                    v[2 * u] = fminf(v[2 * u], sa);
                    v[2 * u + 1] = fminf(v[2 * u + 1], sb);
                }
            }
        }
    }
    long long t1 = clock64();
    float acc = 0;
    for (int u = 0; u < 8; ++u) out[(blockIdx.x * blockDim.x + threadIdx.x) * 8 + u] = v[u];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    float *d_in, *d_o0, *d_o1; long long *d_c;
    float h[2048];
    unsigned s = 12345;
    for (int i = 0; i < 2048; ++i) { s = s * 1664525u + 1013904223u; h[i] = (s >> 8) * (1.0f / 16777216.0f); }
    cudaMalloc(&d_in, sizeof(h)); cudaMemcpy(d_in, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaMalloc(&d_o0, 4 * 8 * 1024); cudaMalloc(&d_o1, 4 * 8 * 1024); cudaMalloc(&d_c, 64);
    static float o0[8 * 1024], o1[8 * 1024];
    for (int w : {1, 4, 7, 8, 14, 16}) {
        for (int nref : {3, 16}) {
            long long c0, c1;
            const int reps = 2000;
            k<0><<<1, 32 * w>>>(d_o0, d_in, nref, reps, d_c, 0x8000000080000000ull); cudaDeviceSynchronize(); cudaMemcpy(&c0, d_c, 8, cudaMemcpyDeviceToHost);
            k<1><<<1, 32 * w>>>(d_o1, d_in, nref, reps, d_c, 0x8000000080000000ull); cudaDeviceSynchronize(); cudaMemcpy(&c1, d_c, 8, cudaMemcpyDeviceToHost);
            cudaMemcpy(o0, d_o0, 4 * 8 * 32 * w, cudaMemcpyDeviceToHost); cudaMemcpy(o1, d_o1, 4 * 8 * 32 * w, cudaMemcpyDeviceToHost);
            int same = 1;
            for (int i = 0; i < 8 * 32 * w; ++i) same &= (*(unsigned *)&o0[i] == *(unsigned *)&o1[i]);
            printf("warps %2d nref %2d: scalar %.1f cyc per (ref x 8 chunks), packed %.1f | bit-equal %d | err %s\n", w, nref,
                   (double)c0 / reps / nref, (double)c1 / reps / nref, same, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
