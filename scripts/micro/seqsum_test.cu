// seqsum_test.cu -- device check of csrc/seqsum.cuh: the tile-parallel evaluation of the strictly sequential binary32 sum
// against the plain dependent-FADD chain, bit for bit, on columns that stress every branch (single-signed, zero-mean,
// exact ties, lattices, mixed magnitudes, infinities, NaN), plus cycles per element of both.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I../../fpsample_b200/csrc -o seqsum_test seqsum_test.cu
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
#include "seqsum.cuh"
using namespace fps;

constexpr int CH = 8192;   // floats staged per chunk
template <int EPL, bool FAST>
__global__ void k(const float *x, u32 n, float *out, long long *cyc, u32 *nfast) {
    __shared__ __align__(16) float buf[CH];
    float sum = 0.0f;
    u32 fast = 0, tiles = 0, hint = 0;
    long long t = 0;
    for (u32 i0 = 0; i0 < n; i0 += CH) {
        const u32 m = min((u32)CH, n - i0);
        for (u32 i = threadIdx.x; i < m; i += 32) buf[i] = x[i0 + i];
        __syncwarp();
        const long long t0 = clock64();
        const u32 a = smem_u32(buf);
        u32 i = 0;
        constexpr u32 TILE = 32 * EPL;
        for (; i + TILE <= m; i += TILE) {
            ++tiles;
            if (FAST && seq_sum_tile<EPL>(a + 4 * i, sum, hint)) ++fast;
            else sum = sq_chain16(a + 4 * i, TILE, sum);
        }
        for (; i < m; ++i) sum = __fadd_rn(sum, buf[i]);
        t += clock64() - t0;
        __syncwarp();
    }
    if (threadIdx.x == 0) *out = sum, *cyc = t, nfast[0] = fast, nfast[1] = tiles;
}

static unsigned rs = 12345;
static float urand() { rs = rs * 1664525u + 1013904223u; return (rs >> 8) * (1.0f / 16777216.0f); }
static float nrand() { float a = urand() + 1e-7f, b = urand(); return sqrtf(-2.f * logf(a)) * cosf(6.2831853f * b); }

int main() {
    const u32 n = 1 << 20;
    std::vector<std::pair<const char *, std::vector<float>>> cols;
    auto add = [&](const char *name, auto gen) { std::vector<float> v(n); for (u32 i = 0; i < n; ++i) v[i] = gen(i); cols.push_back({name, v}); };
    add("uniform [0,1)", [](u32) { return urand(); });
    add("uniform - 0.5", [](u32) { return urand() - 0.5f; });
    add("gauss * 30", [](u32) { return nrand() * 30.f; });
    add("lidar-like x", [](u32) { float az = urand() * 6.2831853f, r = 2.f + 78.f * urand() * urand(); return r * cosf(az); });
    add("negative", [](u32) { return -urand() * 7.f; });
    add("halves (ties)", [](u32) { return (float)((int)(urand() * 100) - 50) * 0.5f; });
    add("quarters>0 (ties)", [](u32) { return (float)((int)(urand() * 64)) * 0.25f; });
    add("lattice 0..5", [](u32) { return (float)((int)(urand() * 6)); });
    add("mixed magnitudes", [](u32 i) { return (i % 97 == 0) ? urand() * 1e6f : urand() * 1e-3f; });
    add("tiny", [](u32) { return urand() * 1e-30f; });
    add("huge", [](u32) { return urand() * 1e30f; });
    add("with inf", [](u32 i) { return i == 700000 ? INFINITY : urand(); });
    add("with nan", [](u32 i) { return i == 300000 ? NAN : urand(); });
    add("sorted ramp", [](u32 i) { return (float)i * 1e-3f - 300.f; });
    add("zeros and ones", [](u32 i) { return (float)(i & 1); });
    float *dx, *dout; long long *dc; u32 *df;
    cudaMalloc(&dx, n * 4); cudaMalloc(&dout, 4); cudaMalloc(&dc, 8); cudaMalloc(&df, 8);
    int bad = 0;
    for (auto &c : cols) {
        cudaMemcpy(dx, c.second.data(), n * 4, cudaMemcpyHostToDevice);
        float ref = 0.f; for (u32 i = 0; i < n; ++i) ref = ref + c.second[i];   // host chain (x86-64: no contraction)
        float o[3]; long long cy[3]; u32 f[3][2];
        k<8, false><<<1, 32>>>(dx, n, dout, dc, df); cudaMemcpy(&o[0], dout, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&cy[0], dc, 8, cudaMemcpyDeviceToHost);
        k<8, true><<<1, 32>>>(dx, n, dout, dc, df); cudaMemcpy(&o[1], dout, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&cy[1], dc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(f[1], df, 8, cudaMemcpyDeviceToHost);
        k<16, true><<<1, 32>>>(dx, n, dout, dc, df); cudaMemcpy(&o[2], dout, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&cy[2], dc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(f[2], df, 8, cudaMemcpyDeviceToHost);
        const bool e0 = !memcmp(&o[0], &ref, 4) || (std::isnan(o[0]) && std::isnan(ref)), e1 = !memcmp(&o[1], &o[0], 4), e2 = !memcmp(&o[2], &o[0], 4);
        bad += !(e0 && e1 && e2);
        printf("%-18s chain==host %d | tile256==chain %d fast %5.1f%% %.2f cyc/elem | tile512==chain %d fast %5.1f%% %.2f cyc/elem | chain %.2f cyc/elem | sum %g  err %s\n",
               c.first, e0, e1, 100.0 * f[1][0] / f[1][1], (double)cy[1] / n, e2, 100.0 * f[2][0] / f[2][1], (double)cy[2] / n, (double)cy[0] / n, o[0],
               cudaGetErrorString(cudaGetLastError()));
    }
    printf(bad ? "MISMATCHES: %d\n" : "all bit-equal (%d)\n", bad);
    return bad != 0;
}
