// common.cuh -- device helpers shared by the FPS kernels (sm_100a only).
//
// Arithmetic contract (reference: src/lib.cpp:214-219, src/_ext/Point.h:48-53, src/_ext/utils.h:12-16):
// squared distances are sums over dimensions in index order of individually rounded binary32
// sub / mul / add.  The _rn intrinsics below are never contracted into FMA by nvcc.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "fpsample_b200 kernels are written for sm_100a (B200) only"
#endif

namespace fps {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr u32 FULL = 0xffffffffu;

// ---- exact arithmetic --------------------------------------------------------------------------
template <int DIM>
__device__ __forceinline__ float sqdist(const float (&p)[DIM], const float (&q)[DIM]) {
    float t = __fsub_rn(p[0], q[0]);
    float acc = __fmul_rn(t, t);  // 0.0f + t*t == t*t exactly
#pragma unroll
    for (int j = 1; j < DIM; ++j) {
        t = __fsub_rn(p[j], q[j]);
        acc = __fadd_rn(acc, __fmul_rn(t, t));
    }
    return acc;
}

// ---- packed binary32 arithmetic (Blackwell add / sub / fma .f32x2 -> FADD2 / FFMA2): two points per instruction -------
// Each half is an individually rounded IEEE operation, so the reference's arithmetic order is kept bit for bit.  ptxas
// contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under --fmad=false, so the product is written fma(t, t, -0.0)
// with the -0.0 pair coming in as a kernel argument (opaque to the compiler): one rounding of t * t, adding -0 changes
// nothing (+0 + -0 = +0), and an FMA result cannot be contracted into the following add.  scripts/micro/f32x2.cu.
__device__ __forceinline__ u64 pk2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void up2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 sq2(u64 t, u64 nz) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(r) : "l"(t), "l"(nz));
    return r;
}
template <int DIM>
__device__ __forceinline__ u64 sqdist2(const u64 (&P)[DIM], const u64 (&Q)[DIM], u64 nz) {
    u64 acc = sq2(sub2(P[0], Q[0]), nz);
#pragma unroll
    for (int j = 1; j < DIM; ++j) acc = add2(acc, sq2(sub2(P[j], Q[j]), nz));
    return acc;
}

// point -> box squared distance, dims in order (reference: src/_ext/KDNode.h:105-118)
template <int DIM>
__device__ __forceinline__ float boxdist(const float (&r)[DIM], const float (&lo)[DIM],
                                         const float (&hi)[DIM]) {
    float acc = 0.0f;
#pragma unroll
    for (int j = 0; j < DIM; ++j) {
        float e = 0.0f;
        if (r[j] > hi[j])
            e = __fsub_rn(r[j], hi[j]);
        else if (r[j] < lo[j])
            e = __fsub_rn(lo[j], r[j]);
        acc = __fadd_rn(acc, __fmul_rn(e, e));
    }
    return acc;
}

// ---- order-preserving float <-> int (for min/max reductions over signed floats) -------------------
__device__ __forceinline__ int f2ord(float f) {
    int b = __float_as_int(f);
    return b ^ ((b >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ord2f(int o) { return __int_as_float(o ^ ((o >> 31) & 0x7fffffff)); }

// ---- 64-bit max-key reductions --------------------------------------------------------------------
// key = (float_bits(dist) << 32) | tiebreak, dist >= 0 so the bits order as unsigned.
__device__ __forceinline__ u64 make_key(float d, u32 low) {
    return ((u64)__float_as_uint(d) << 32) | (u64)low;
}

__device__ __forceinline__ u64 warp_max_key(u64 key) {
    u32 hi = (u32)(key >> 32), lo = (u32)key;
    u32 mhi = __reduce_max_sync(FULL, hi);
    u32 cand = (hi == mhi) ? lo : 0u;
    u32 mlo = __reduce_max_sync(FULL, cand);
    return ((u64)mhi << 32) | (u64)mlo;
}

__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ u32 warp_id() { return threadIdx.x >> 5; }

// ---- cluster / DSMEM / mbarrier PTX ----------------------------------------------------------------
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ u32 cluster_ctarank() {
    u32 r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ u32 cluster_nctarank() {
    u32 r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// map a local shared address to the same offset in CTA `rank` of the cluster
__device__ __forceinline__ u32 mapa(u32 local_smem_addr, u32 rank) {
    u32 r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_u64(u32 addr, u64 v) {
    asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_f32(u32 addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void mbar_init(u32 bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init_cluster() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// arrive (release, cluster scope) on an mbarrier that may live in another CTA of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(u32 cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(u32 bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(u32 bar, u32 parity) {
    u32 ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(u32 bar, u32 parity) {
    while (!mbar_try_wait_cluster(bar, parity)) {
    }
}
// TMA 1-D bulk copy global -> shared::cta, completion counted in bytes on an mbarrier.
// Requires 16-byte aligned source, destination and size.
__device__ __forceinline__ void tma_bulk_g2s(u32 dst_smem, const void *src, u32 bytes, u32 bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

}  // namespace fps
