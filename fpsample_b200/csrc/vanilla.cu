// vanilla.cu -- exact (brute-force) farthest point sampling, reference semantics of
// fps_sampling / fps_sampling_multi_start_index (src/lib.cpp:188-246, 111-186):
//   dist_min = +inf; K-1 rounds of { min-update against the previous pick ; argmax with the HIGHEST
//   index among maxima }.  Forced start picks still min-update.
//
// Two kernels:
//   vanilla_cluster_kernel<DIM,PPT>  one thread-block cluster (1..16 CTAs) per cloud.  Coordinates and
//       running min-distances live in REGISTERS for all K rounds; a row-major copy of the CTA's slice
//       stays in shared memory (staged once with a TMA bulk copy) only to look the winner's
//       coordinates up.  Argmax = 64-bit key (dist bits, index) reduced with redux.sync inside a warp,
//       one __syncthreads per round inside a CTA, and a DSMEM store + remote mbarrier arrive per
//       round across the cluster (no cluster-wide barrier, no global memory on the critical path).
//   vanilla_grid_kernel               any n, any dim: cooperative groups of CTAs, coordinates in shared
//       memory (if they fit) or L2/HBM (SoA), one global-memory barrier per round per group.
#include "common.cuh"
#include "engine.h"

namespace fps {

// ======================================================================================================
//  cluster kernel
// ======================================================================================================
constexpr int VT = 512;          // threads per CTA
constexpr int VNW = VT / 32;     // warps per CTA
constexpr int MAXC = 16;         // max cluster size (non-portable)

struct XSlot {                   // one CTA's candidate, written into every CTA of the cluster
    u64 key;
    float c[8];
};

struct VanillaSmem {
    u64 xbar;                    // mbarrier: cluster exchange
    u64 tbar;                    // mbarrier: TMA staging
    u64 wslot[2][VNW];           // per-warp candidates, double buffered
    XSlot xslot[2][MAXC];        // per-CTA candidates, double buffered
};

template <int DIM, int PPT>
__global__ void __launch_bounds__(VT, (PPT * (DIM + 1) <= 36) ? 2 : 1)
vanilla_cluster_kernel(VanillaArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    VanillaSmem &S = *reinterpret_cast<VanillaSmem *>(smem_raw);
    float *scoord = reinterpret_cast<float *>(smem_raw + ((sizeof(VanillaSmem) + 15) & ~15));

    const u32 C = cluster_nctarank();
    const u32 rank = cluster_ctarank();
    const u32 cloud = blockIdx.x / C;
    const u32 tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    const u32 n = a.n, dim = a.dim;
    const u32 lo = rank * a.slice;
    const u32 cnt = lo < n ? min(a.slice, n - lo) : 0u;
    const float *gcloud = a.pts + (size_t)cloud * n * dim;
    const float *gsrc = gcloud + (size_t)lo * dim;

    // ---- stage the slice (row-major) into shared memory: TMA bulk copy + scalar tail ------------------
    const u32 total = cnt * dim;
    const bool aligned = ((reinterpret_cast<uintptr_t>(gsrc) & 15) == 0);
    const u32 bulk = aligned ? (total & ~3u) : 0u;
    if (tid == 0) {
        mbar_init(smem_u32(&S.xbar), C);
        mbar_init(smem_u32(&S.tbar), 1);
        fence_mbar_init_cluster();
    }
    __syncthreads();
    if (bulk && tid == 0) {
        mbar_arrive_expect_tx(smem_u32(&S.tbar), bulk * 4);
        // chunks keep each transaction well inside the mbarrier tx-count range
        for (u32 off = 0; off < bulk; off += 16384) {
            u32 len = min(16384u, bulk - off);
            tma_bulk_g2s(smem_u32(scoord + off), gsrc + off, len * 4, smem_u32(&S.tbar));
        }
    }
    for (u32 i = bulk + tid; i < total; i += VT) scoord[i] = gsrc[i];
    if (bulk) mbar_wait_cluster(smem_u32(&S.tbar), 0);
    __syncthreads();
    if (C > 1) cluster_sync_all();  // every CTA's exchange mbarrier is initialised before remote arrives

    // ---- registers: my points and their running min distances ----------------------------------------
    static_assert(PPT % 2 == 0, "points are held in pairs");
    u64 P2[PPT / 2][DIM];   // points (2j, 2j + 1) of this thread, packed per dimension
    float dm[PPT];
#pragma unroll
    for (int j = 0; j < PPT; j += 2) {
        const u32 l0 = j * VT + tid, l1 = (j + 1) * VT + tid;
        const bool v0 = l0 < cnt, v1 = l1 < cnt;
#pragma unroll
        for (int c = 0; c < DIM; ++c)
            P2[j / 2][c] = pk2((v0 && c < dim) ? scoord[l0 * dim + c] : 0.0f, (v1 && c < dim) ? scoord[l1 * dim + c] : 0.0f);
        dm[j] = v0 ? __int_as_float(0x7f800000) : -1.0f;      // +inf (lib.cpp:206); -1 never wins
        dm[j + 1] = v1 ? __int_as_float(0x7f800000) : -1.0f;
    }
    const u64 nz = a.negzero;

    const u32 k = a.k, ns = a.n_starts;
    const u64 *starts = a.starts ? a.starts + (size_t)cloud * ns : nullptr;
    u32 cur = starts ? (u32)starts[0] : 0u;
    float q[DIM];
#pragma unroll
    for (int c = 0; c < DIM; ++c) q[c] = (c < dim) ? gcloud[(size_t)cur * dim + c] : 0.0f;
    u64 *out = a.out + (size_t)cloud * k;
    const bool writer = (rank == 0 && tid == 0);
    if (writer) out[0] = cur;

    for (u32 t = 1; t < k; ++t) {
        const u32 par = (t - 1) & 1;
        // min-update + thread-local argmax, ascending index order so '>=' keeps the highest index
        float best = -1.0f;
        u32 bj = 0;
        u64 Q2[DIM];
#pragma unroll
        for (int c = 0; c < DIM; ++c) Q2[c] = pk2(q[c], q[c]);
#pragma unroll
        for (int j = 0; j < PPT; j += 2) {
            float d0, d1;
            up2(sqdist2<DIM>(P2[j / 2], Q2, nz), d0, d1);
            const float v0 = fminf(dm[j], d0), v1 = fminf(dm[j + 1], d1);
            dm[j] = v0, dm[j + 1] = v1;
            if (v0 >= best) best = v0, bj = j;
            if (v1 >= best) best = v1, bj = j + 1;
        }
        u64 key = (best < 0.0f) ? 0ull : make_key(best, lo + bj * VT + tid);
        key = warp_max_key(key);
        if (lane == 0) S.wslot[par][warp] = key;
        __syncthreads();
        u64 k2 = (lane < VNW) ? S.wslot[par][lane] : 0ull;
        k2 = warp_max_key(k2);  // CTA winner, known to every thread

        if (C == 1) {
            cur = (u32)k2;
#pragma unroll
            for (int c = 0; c < DIM; ++c) q[c] = (c < dim) ? scoord[cur * dim + c] : 0.0f;
        } else {
            if (warp == 0 && lane < C) {
                const u32 dst = mapa(smem_u32(&S.xslot[par][rank]), lane);
                st_cluster_u64(dst, k2);
                if (cnt) {
                    const u32 lw = (u32)k2 - lo;
                    for (u32 c = 0; c < dim; ++c) st_cluster_f32(dst + 8 + 4 * c, scoord[lw * dim + c]);
                }
                mbar_arrive_cluster(mapa(smem_u32(&S.xbar), lane));
            }
            mbar_wait_cluster(smem_u32(&S.xbar), par);
            u64 mine = (lane < C) ? S.xslot[par][lane].key : 0ull;
            u64 k3 = warp_max_key(mine);
            const u32 wr = __ffs(__ballot_sync(FULL, lane < C && mine == k3)) - 1;
            cur = (u32)k3;
#pragma unroll
            for (int c = 0; c < DIM; ++c) q[c] = (c < dim) ? S.xslot[par][wr].c[c] : 0.0f;
        }
        if (t < ns) {  // forced pick (lib.cpp:151-155): overrides the argmax, rare
            cur = (u32)starts[t];
#pragma unroll
            for (int c = 0; c < DIM; ++c) q[c] = (c < dim) ? gcloud[(size_t)cur * dim + c] : 0.0f;
        }
        if (writer) out[t] = cur;
    }
    if (C > 1) cluster_sync_all();  // nobody exits while a peer may still touch its shared memory
}

// ======================================================================================================
//  grid kernel (generic fallback): groups of G CTAs, one cloud per group at a time
// ======================================================================================================
constexpr int GT = 512;
constexpr int GNW = GT / 32;

__device__ __forceinline__ u32 ld_acquire_u32(const u32 *p) {
    u32 v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ u64 ld_relaxed_u64(const u64 *p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// group barrier: monotonically increasing arrival counter, target = generation * G
__device__ __forceinline__ void group_barrier(u32 *counter, u32 target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        while (ld_acquire_u32(counter) < target) {
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(GT, 1) vanilla_grid_kernel(VanillaGridArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ u64 wslot[GNW];
    const u32 G = a.G;
    const u32 group = blockIdx.x / G, grank = blockIdx.x % G, ngroups = gridDim.x / G;
    const u32 tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    const u32 n = a.n, dim = a.dim, k = a.k, ns = a.n_starts;
    const u32 slice = a.slice;  // points per CTA
    const u32 lo = grank * slice;
    const u32 cnt = lo < n ? min(slice, n - lo) : 0u;
    u32 *counter = a.counters + group * 32;      // one 128-byte line per group
    u64 *gslots = a.slots + (size_t)group * 2 * G;  // [2][G]
    u32 gen = 0;

    for (u32 cloud = group; cloud < a.B; cloud += ngroups) {
        const float *gcloud = a.pts + (size_t)cloud * n * dim;
        // my slice: coordinates SoA [dim][slice] + dm[slice], in shared memory or in the workspace
        float *co, *dmv;
        if (a.use_smem) {
            co = reinterpret_cast<float *>(smem_raw);
        } else {
            co = a.scratch + ((size_t)group * G + grank) * (size_t)slice * (dim + 1);
        }
        dmv = co + (size_t)slice * dim;
        for (u32 i = tid; i < cnt * dim; i += GT) {
            u32 pt = i / dim, c = i - pt * dim;
            co[(size_t)c * slice + pt] = gcloud[(size_t)lo * dim + i];
        }
        for (u32 i = tid; i < cnt; i += GT) dmv[i] = __int_as_float(0x7f800000);
        __syncthreads();

        const u64 *starts = a.starts ? a.starts + (size_t)cloud * ns : nullptr;
        u32 cur = starts ? (u32)starts[0] : 0u;
        u64 *out = a.out + (size_t)cloud * k;
        const bool writer = (grank == 0 && tid == 0);
        if (writer) out[0] = cur;

        for (u32 t = 1; t < k; ++t) {
            const u32 par = t & 1;
            const float *qrow = gcloud + (size_t)cur * dim;
            u64 key = 0;
            for (u32 i = tid; i < cnt; i += GT) {
                float t0 = __fsub_rn(co[i], __ldg(qrow));
                float acc = __fmul_rn(t0, t0);
                for (u32 c = 1; c < dim; ++c) {
                    float tc = __fsub_rn(co[(size_t)c * slice + i], __ldg(qrow + c));
                    acc = __fadd_rn(acc, __fmul_rn(tc, tc));
                }
                float v = fminf(dmv[i], acc);
                dmv[i] = v;
                u64 kk = make_key(v, lo + i);
                key = kk > key ? kk : key;  // ascending i: equal dist -> larger index -> larger key
            }
            key = warp_max_key(key);
            if (lane == 0) wslot[warp] = key;
            __syncthreads();
            if (warp == 0) {
                u64 k2 = (lane < GNW) ? wslot[lane] : 0ull;
                k2 = warp_max_key(k2);
                if (lane == 0) gslots[par * G + grank] = k2;
            }
            ++gen;
            group_barrier(counter, gen * G);
            u64 best = 0;
            for (u32 i = lane; i < G; i += 32) {
                u64 v = ld_relaxed_u64(gslots + par * G + i);
                best = v > best ? v : best;
            }
            best = warp_max_key(best);
            cur = (u32)best;
            if (t < ns) cur = (u32)starts[t];
            if (writer) out[t] = cur;
        }
        __syncthreads();
    }
}

// ======================================================================================================
//  host side: plan + launch
// ======================================================================================================
template <int DIM, int PPT>
static cudaError_t launch_cluster(const VanillaArgs &a, u32 B, u32 C, size_t smem, cudaStream_t st) {
    auto kern = vanilla_cluster_kernel<DIM, PPT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (C > 8) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(B * C);
    cfg.blockDim = dim3(VT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, a);
}

template <int DIM>
static cudaError_t dispatch_ppt(int ppt, const VanillaArgs &a, u32 B, u32 C, size_t smem, cudaStream_t st) {
    switch (ppt) {
        case 2: return launch_cluster<DIM, 2>(a, B, C, smem, st);
        case 4: return launch_cluster<DIM, 4>(a, B, C, smem, st);
        case 8: return launch_cluster<DIM, 8>(a, B, C, smem, st);
        case 16:
            if constexpr (DIM <= 4) return launch_cluster<DIM, 16>(a, B, C, smem, st);
            break;
        case 24:
            if constexpr (DIM <= 3) return launch_cluster<DIM, 24>(a, B, C, smem, st);
            break;
    }
    return cudaErrorInvalidValue;
}

static int max_ppt(int dimp) { return dimp <= 3 ? 24 : (dimp <= 4 ? 16 : 8); }
static int pad_dim(int dim) { return dim <= 2 ? 2 : dim == 3 ? 3 : dim == 4 ? 4 : dim <= 6 ? 6 : 8; }

// Decide how a vanilla batch is run.  Returns false if only the grid kernel can take it.
bool plan_vanilla_cluster(size_t n, size_t dim, size_t B, int n_sms, VanillaPlan *pl) {
    if (dim == 0 || dim > 8 || n == 0) return false;
    const int dimp = pad_dim((int)dim);
    const size_t cap = (size_t)VT * max_ppt(dimp);
    size_t C = (n + cap - 1) / cap;
    if (C > MAXC) return false;
    while (C & (C - 1)) ++C;  // cluster sizes 1,2,4,8,16 only
    // few clouds: spread each over more SMs (shorter rounds) while the machine still has room
    while (C < 8 && B * C * 2 <= (size_t)n_sms && (n + C - 1) / C > 8192) C *= 2;
    // slice: multiple of 4 points so each CTA's rows start 16-byte aligned whenever the cloud does
    size_t slice = ((n + C - 1) / C + 3) & ~(size_t)3;
    int ppt = 2;
    for (int cand : {2, 4, 8, 16, 24})
        if ((size_t)cand * VT >= slice) {
            ppt = cand;
            break;
        }
    if ((size_t)ppt * VT < slice || ppt > max_ppt(dimp)) return false;
    size_t smem = ((sizeof(VanillaSmem) + 15) & ~15) + slice * dim * sizeof(float) + 16;
    if (smem > 227 * 1024) return false;
    pl->dimp = dimp;
    pl->ppt = ppt;
    pl->C = (u32)C;
    pl->slice = (u32)slice;
    pl->smem = smem;
    return true;
}

cudaError_t launch_vanilla_cluster(const VanillaPlan &pl, VanillaArgs a, u32 B, cudaStream_t st) {
    a.negzero = 0x8000000080000000ull;
    a.slice = pl.slice;
    switch (pl.dimp) {
        case 2: return dispatch_ppt<2>(pl.ppt, a, B, pl.C, pl.smem, st);
        case 3: return dispatch_ppt<3>(pl.ppt, a, B, pl.C, pl.smem, st);
        case 4: return dispatch_ppt<4>(pl.ppt, a, B, pl.C, pl.smem, st);
        case 6: return dispatch_ppt<6>(pl.ppt, a, B, pl.C, pl.smem, st);
        case 8: return dispatch_ppt<8>(pl.ppt, a, B, pl.C, pl.smem, st);
    }
    return cudaErrorInvalidValue;
}

// grid kernel plan: G CTAs per group, as many groups as fit in one co-resident wave
void plan_vanilla_grid(size_t n, size_t dim, size_t B, int n_sms, VanillaGridPlan *pl) {
    // a CTA is comfortable with <= 8192 points per round; use more CTAs per cloud for big clouds
    size_t G = (n + 8191) / 8192;
    if (G < 1) G = 1;
    if (G > (size_t)n_sms) G = n_sms;
    size_t groups = (size_t)n_sms / G;
    if (groups > B) groups = B;
    if (groups < 1) groups = 1;
    if (groups == 1 || B == 1) {  // single cloud: whole machine
        G = (size_t)n_sms;
        if (G * 64 > n) G = (n + 63) / 64;
        groups = 1;
    }
    size_t slice = ((n + G - 1) / G + 3) & ~(size_t)3;
    size_t bytes = slice * (dim + 1) * sizeof(float);
    pl->G = (u32)G;
    pl->groups = (u32)groups;
    pl->slice = (u32)slice;
    pl->use_smem = bytes <= 200 * 1024;
    pl->smem = pl->use_smem ? bytes : 0;
    pl->scratch_floats = pl->use_smem ? 0 : groups * G * slice * (dim + 1);
}

cudaError_t launch_vanilla_grid(const VanillaGridPlan &pl, VanillaGridArgs a, cudaStream_t st) {
    a.G = pl.G;
    a.slice = pl.slice;
    a.use_smem = pl.use_smem ? 1 : 0;
    cudaError_t e = cudaFuncSetAttribute(vanilla_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)pl.smem);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(a.counters, 0, (size_t)pl.groups * 32 * sizeof(u32), st);
    if (e != cudaSuccess) return e;
    void *params[] = {&a};
    return cudaLaunchCooperativeKernel((void *)vanilla_grid_kernel, dim3(pl.G * pl.groups), dim3(GT), params,
                                       pl.smem, st);
}

}  // namespace fps
