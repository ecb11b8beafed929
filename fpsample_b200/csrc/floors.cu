// floors.cu -- latency floors for the pick-latency-bound configurations (SURVEY.md 8(d) "Which roofline governs": single
// clouds and small batches are K-1 DEPENDENT picks, so the bound is the cost of one empty round of the sampler's own
// synchronisation structure, not bytes or flops).  Each kernel below runs `rounds` rounds of exactly the exchange its
// sampler performs per round -- same instructions, same scopes, same message sizes -- with no distance work in between;
// bench.py times it with CUDA events and reports  floor x rounds / kernel time  as roofline.frac with bound = "latency".
//
//   FPS_FLOOR_WARP    one warp: redux.max + redux.min + ballot + 3 shuffles + one shared load
//                     (the arg-max of kdline_warp_kernel, csrc/kdline_warp.cu)
//   FPS_FLOOR_CLUSTER cluster of C CTAs x 512 threads: warp redux -> shared slot -> __syncthreads -> CTA redux -> DSMEM store of
//                     (key, coordinates) into every peer + remote mbarrier arrive -> wait -> redux over the C slots
//                     (the round of vanilla_cluster_kernel, csrc/vanilla.cu)
//   FPS_FLOOR_GRID    G CTAs x 1024 threads (cooperative): every CTA publishes `words` stamped 16-byte words in L2, every CTA
//                     gathers all G x words of them by polling the stamps, two __syncthreads and a redux around it
//                     (the round of kdline_grid_kernel, csrc/kdline_grid.cu)
#include "../../include/fps_b200.h"
#include "common.cuh"
#include "engine.h"

namespace fps {

__global__ void __launch_bounds__(32, 1) floor_warp_kernel(u32 rounds, u32 *sink) {
    __shared__ float pts[3 * 128];
    const u32 lane = lane_id();
    for (u32 i = lane; i < 3 * 128; i += 32) pts[i] = (float)i;
    __syncwarp();
    u32 key = lane * 2654435761u, pos = lane;
    float acc = 0.f;
    for (u32 r = 0; r < rounds; ++r) {
        const u32 M = __reduce_max_sync(FULL, key);
        const u32 mine = (key == M) ? pos : 0xffffffffu;
        const u32 cur = __reduce_min_sync(FULL, mine);
        const u32 src = __ffs(__ballot_sync(FULL, mine == cur)) - 1;
        float c0 = __shfl_sync(FULL, (float)key, src), c1 = __shfl_sync(FULL, (float)pos, src), c2 = __shfl_sync(FULL, acc, src);
        acc += pts[(cur + r) & 127] + c0 + c1 + c2;
        key = key * 1664525u + 1013904223u + (u32)acc;   // the next round depends on this one
        pos = (pos + cur + 1) & 0xffffu;
    }
    sink[lane] = key + (u32)acc;
}

struct XSlot {
    u64 key;
    float c[8];
};

__global__ void __launch_bounds__(512, 1) floor_cluster_kernel(u32 rounds, u32 *sink) {
    __shared__ u64 wslot[2][16];
    __shared__ __align__(16) XSlot xslot[2][16];
    __shared__ u64 xbar;
    const u32 tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    const u32 C = cluster_nctarank(), rank = cluster_ctarank();
    if (tid == 0) {
        mbar_init(smem_u32(&xbar), C);
        fence_mbar_init_cluster();
    }
    __syncthreads();
    if (C > 1) cluster_sync_all();
    u64 key = ((u64)(tid * 2654435761u) << 32) | (rank * 512 + tid);
    for (u32 t = 1; t <= rounds; ++t) {
        const u32 par = (t - 1) & 1;
        u64 k1 = warp_max_key(key);
        if (lane == 0) wslot[par][warp] = k1;
        __syncthreads();
        u64 k2 = (lane < 16) ? wslot[par][lane] : 0ull;
        k2 = warp_max_key(k2);
        if (C > 1) {
            if (warp == 0 && lane < C) {
                const u32 dst = mapa(smem_u32(&xslot[par][rank]), lane);
                st_cluster_u64(dst, k2);
                for (u32 c = 0; c < 3; ++c) st_cluster_f32(dst + 8 + 4 * c, (float)(u32)k2);
                mbar_arrive_cluster(mapa(smem_u32(&xbar), lane));
            }
            mbar_wait_cluster(smem_u32(&xbar), par);
            u64 mine = (lane < C) ? xslot[par][lane].key : 0ull;
            const u64 k3 = warp_max_key(mine);
            const u32 wr = __ffs(__ballot_sync(FULL, lane < C && mine == k3)) - 1;
            k2 = k3 + (u64)xslot[par][wr].c[0];
        }
        key = key * 6364136223846793005ull + k2;   // the next round depends on this one
    }
    if (C > 1) cluster_sync_all();
    sink[blockIdx.x * 512 + tid] = (u32)key;
}

__device__ __forceinline__ uint4 ld_relaxed_v4(const uint4 *p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_v4(uint4 *p, u32 x, u32 y, u32 z, u32 w) {
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

// groups of `gc` CTAs exchange among themselves (gc = gridDim.x: the whole grid, one huge cloud; gc = 2: BASELINE cfg 3)
__global__ void __launch_bounds__(1024, 1) floor_grid_kernel(uint4 *pub, u32 rounds, u32 words, u32 gc, u32 *sink) {
    extern __shared__ uint4 gbuf[];
    const u32 tid = threadIdx.x, grp = blockIdx.x / gc, cta = blockIdx.x % gc;
    uint4 *base = pub + (size_t)grp * 2 * gc * words;
    u32 acc = 0;
    for (u32 round = 0; round < rounds; ++round) {
        const u32 stamp = round + 1;
        uint4 *dst = base + ((size_t)(round & 1) * gc + cta) * words;
        if (tid < words) st_relaxed_v4(dst + tid, stamp, tid + acc, cta, round);
        for (u32 i = tid; i < gc * words; i += 1024) {
            const uint4 *src = base + (size_t)(round & 1) * gc * words + i;
            uint4 v;
            do { v = ld_relaxed_v4(src); } while (v.x != stamp);
            gbuf[i] = v;
        }
        __syncthreads();
        u64 b = (tid < gc * words) ? (((u64)gbuf[tid].y << 32) | gbuf[tid].z) : 0ull;
        b = warp_max_key(b);
        if ((tid & 31) == 0) reinterpret_cast<u64 *>(gbuf + gc * words)[tid >> 5] = b;
        __syncthreads();
        acc += (u32)reinterpret_cast<u64 *>(gbuf + gc * words)[tid & 3];   // the next round depends on this one
    }
    sink[blockIdx.x * 1024 + tid] = acc;
}

}  // namespace fps

using namespace fps;

extern "C" int fps_b200_sync_floor(int kind, int ctas, int words, int rounds, float *ns_per_round) {
    if (!ns_per_round || rounds < 1 || ctas < 1) return FPS_ERR_ARG;
    u32 *sink = nullptr;
    uint4 *pub = nullptr;
    cudaEvent_t e0, e1;
    cudaError_t e = cudaMalloc(&sink, (size_t)148 * 1024 * 4 * 2);
    if (e != cudaSuccess) return FPS_ERR_CUDA + (int)e;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    int rc = FPS_OK;
    for (int rep = 0; rep < 3 && rc == FPS_OK; ++rep) {   // the first repetition also warms the instruction cache
        if (kind == FPS_FLOOR_WARP) {
            cudaEventRecord(e0);
            floor_warp_kernel<<<1, 32>>>((u32)rounds, sink);
            cudaEventRecord(e1);
        } else if (kind == FPS_FLOOR_CLUSTER) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)ctas);
            cfg.blockDim = dim3(512);
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = (unsigned)ctas;
            at[0].val.clusterDim.y = at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            if (ctas > 8) cudaFuncSetAttribute(floor_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            cudaEventRecord(e0);
            e = cudaLaunchKernelEx(&cfg, floor_cluster_kernel, (u32)rounds, sink);
            cudaEventRecord(e1);
        } else if (kind == FPS_FLOOR_GRID) {
            const int gc = words >> 16 ? words >> 16 : ctas;   // words = (CTAs per group << 16) | 16-byte words per CTA
            const int w = words & 0xffff;
            if (w < 1 || w > 1024 || gc < 1 || ctas % gc) {
                rc = FPS_ERR_ARG;
                break;
            }
            const size_t pub_bytes = (size_t)2 * ctas * w * 16, smem = (size_t)gc * w * 16 + 256;
            if (!pub && (e = cudaMalloc(&pub, pub_bytes)) != cudaSuccess) break;
            cudaMemset(pub, 0, pub_bytes);
            cudaFuncSetAttribute(floor_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            u32 r = (u32)rounds, ww = (u32)w, g = (u32)gc;
            void *args[] = {&pub, &r, &ww, &g, &sink};
            cudaEventRecord(e0);
            e = cudaLaunchCooperativeKernel((void *)floor_grid_kernel, dim3((unsigned)ctas), dim3(1024), args, smem, 0);
            cudaEventRecord(e1);
        } else {
            rc = FPS_ERR_ARG;
            break;
        }
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
        count_launch();
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    if (pub) cudaFree(pub);
    if (rc) return rc;
    if (e != cudaSuccess) return FPS_ERR_CUDA + (int)e;
    *ns_per_round = best * 1e6f / (float)rounds;
    return FPS_OK;
}
