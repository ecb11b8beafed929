// kdbuild.cu -- kd-line BUILD with the whole grid per tree level, for clouds (or small batches of clouds) that one
// CTA per cloud would leave most of the machine idle on.  Same bit-exact result as kdline.cu's build, i.e. the
// permutation, leaf ranges and tight leaf boxes of the reference's recursive divideTree
// (src/_ext/KDTreeBase.h:84-207, leaf rule src/_ext/KDLineTree.h:37-39), written into the per-cloud region the
// asynchronous sampler reads:  [q dim*npad f32][scratch/dis npad][perm npad u32][nlo pad32(S+1) u32][box S*2*dim].
//
// Per level l (node j of level l lives at slot j << (h-l); children reuse slot / slot+half):
//   gb_split   one warp per node: split dim (first max span, KDTreeBase.h:160-179) + SEQUENTIAL binary32 mean
//              (KDTreeBase.h:151-158; the dependent add chain is the only serial piece: 4 cycles per point)
//   gb_items   one warp per cloud: work items (1024 positions of one node each) -> item table
//   gb_count   one warp per item: points '< split value'
//   gb_scan    one warp per node: exclusive prefix of the item counts, m = total
//   gb_rank    one warp per item: rank the misplaced points of both sides (closed form of the Hoare loop,
//              KDTreeBase.h:123-149: k-th misplaced from the left pairs with the k-th from the right)
//   gb_swap    one warp per item: the swaps; child boundaries; child box reset
//   gb_box     one warp per item: tight child boxes (KDTreeBase.h:181-207) by redux + atomics on ordered ints
#include <algorithm>
#include <cfloat>

#include "common.cuh"
#include "engine.h"
#include "kdcommon.cuh"

namespace fps {

constexpr u32 GB_WCH = 1024;   // positions per work item (one warp)
constexpr u32 GB_WPB = 8;      // warps per CTA in the item kernels

struct GBArgs {
    const float *pts;
    unsigned char *region;
    size_t region_stride;
    unsigned char *aux;   // per cloud: A0[S] A1[S] A2[S] A3[S] nitems[S] ibase[S+1 -> pad] part[wmax]
    size_t aux_stride;
    u32 B, n, npad, dim, h, S, nlo_pad, wmax, lvl;
};

struct GBView {
    float *q;
    u32 *scr, *perm, *nlo;
    int *box;
    u32 *A0, *A1, *A2, *A3, *nitems, *ibase, *part;
};

__device__ __forceinline__ GBView gb_view(const GBArgs &a, u32 cloud) {
    GBView v;
    unsigned char *rg = a.region + (size_t)cloud * a.region_stride;
    v.q = reinterpret_cast<float *>(rg);
    v.scr = reinterpret_cast<u32 *>(rg) + (size_t)a.dim * a.npad;
    v.perm = v.scr + a.npad;
    v.nlo = v.perm + a.npad;
    v.box = reinterpret_cast<int *>(v.nlo + a.nlo_pad);
    u32 *ax = reinterpret_cast<u32 *>(a.aux + (size_t)cloud * a.aux_stride);
    v.A0 = ax;
    v.A1 = ax + a.S;
    v.A2 = ax + 2 * a.S;
    v.A3 = ax + 3 * a.S;
    v.nitems = ax + 4 * a.S;
    v.ibase = ax + 5 * a.S;
    v.part = ax + 6 * a.S + 32;
    return v;
}

// ---- stage: row-major -> SoA, identity permutation, slot table, root box reset ---------------------------------
__global__ void __launch_bounds__(256) gb_stage(GBArgs a) {
    const u32 cloud = blockIdx.y;
    GBView v = gb_view(a, cloud);
    const float *g = a.pts + (size_t)cloud * a.n * a.dim;
    const u32 total = a.n * a.dim;
    for (u32 f = blockIdx.x * blockDim.x + threadIdx.x; f < total; f += gridDim.x * blockDim.x) {
        const u32 i = f / a.dim, c = f - i * a.dim;
        v.q[(size_t)c * a.npad + i] = g[f];
        if (c == 0) v.perm[i] = i;
    }
    if (blockIdx.x == 0) {
        for (u32 s = threadIdx.x; s <= a.S; s += blockDim.x) v.nlo[s] = (s == a.S) ? a.n : 0u;
        if (threadIdx.x < 2 * a.dim) v.box[threadIdx.x] = (threadIdx.x < a.dim) ? 0x7fffffff : (int)0x80000000;
    }
}

// root box: every warp takes 1024 positions (all "left of the split")
template <int DIM>
__global__ void __launch_bounds__(256) gb_rootbox(GBArgs a) {
    const u32 cloud = blockIdx.y;
    GBView v = gb_view(a, cloud);
    const u32 w = blockIdx.x * (blockDim.x >> 5) + warp_id();
    const u32 s0 = w * GB_WCH;
    if (s0 >= a.n) return;
    const u32 s1 = min(a.n, s0 + GB_WCH);
    box_range<DIM>(v.q, a.npad, a.dim, s0, s1, a.n, v.box, v.box);
}

// ---- gb_split: one warp per node ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) gb_split(GBArgs a) {
    __shared__ __align__(16) float ring[4][SS_STAGES * SS_TILE];   // one TMA ring per warp
    __shared__ u64 bars[4][SS_STAGES];
    if (lane_id() < SS_STAGES) mbar_init(smem_u32(&bars[warp_id()][lane_id()]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    u32 phase = 0;
    const u32 nn = 1u << a.lvl, stride = a.S >> a.lvl, half = stride >> 1;
    const u32 gw = blockIdx.x * (blockDim.x >> 5) + warp_id();
    if (gw >= a.B * nn) return;
    const u32 cloud = gw / nn, j = gw - cloud * nn;
    GBView v = gb_view(a, cloud);
    const u32 lane = lane_id();
    const u32 idx = j * stride;
    const u32 lo = v.nlo[idx], hi = v.nlo[idx + stride], count = hi - lo;
    if (count < 2) {  // no split: everything stays in the left child, the right child is an empty slot
        if (lane == 0) {
            v.nlo[idx + half] = hi;
            v.nitems[j] = 0;
        }
        if (lane < 2 * a.dim) v.box[(size_t)(idx + half) * 2 * a.dim + lane] = (lane < a.dim) ? 0x7fffffff : (int)0x80000000;
        return;
    }
    const int *b = v.box + (size_t)idx * 2 * a.dim;
    u32 sd = 0;
    float span = 0.0f;
    for (u32 c = 0; c < a.dim; ++c) {
        const float s = __fsub_rn(ord2f(b[a.dim + c]), ord2f(b[c]));
        if (s > span) {
            span = s;
            sd = c;
        }
    }
    const float sum = seq_sum_tma(v.q + (size_t)sd * a.npad + lo, count, ring[warp_id()], bars[warp_id()], phase);
    const float val = __fdiv_rn(sum, __uint2float_rn(count));
    if (lane == 0) {
        v.A0[j] = __float_as_uint(val);
        v.A1[j] = sd;
        v.A3[j] = 0;
        v.nitems[j] = (count + GB_WCH - 1) / GB_WCH;
    }
}

// ---- gb_items: one warp per cloud, exclusive prefix of the nodes' item counts ----------------------------------------
__global__ void __launch_bounds__(32) gb_items(GBArgs a) {
    const u32 cloud = blockIdx.x, nn = 1u << a.lvl, lane = lane_id();
    GBView v = gb_view(a, cloud);
    u32 run = 0;
    for (u32 j0 = 0; j0 < nn; j0 += 32) {
        const u32 j = j0 + lane;
        const u32 x = j < nn ? v.nitems[j] : 0u;
        u32 inc = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 y = __shfl_up_sync(FULL, inc, o);
            if ((int)lane >= o) inc += y;
        }
        if (j < nn) v.ibase[j] = run + inc - x;
        run += __shfl_sync(FULL, inc, 31);
    }
    if (lane == 0) v.ibase[nn] = run;
}

// which (node, sub-range) a warp of the item kernels works on; false if it has nothing to do
struct GBItem {
    u32 cloud, j, r, idx, lo, hi;
};
__device__ __forceinline__ bool gb_item(const GBArgs &a, const GBView &v, u32 cloud, u32 wi, GBItem *it) {
    const u32 nn = 1u << a.lvl, stride = a.S >> a.lvl;
    if (wi >= v.ibase[nn]) return false;
    u32 l = 0, r = nn;  // last node with ibase[j] <= wi
    while (r - l > 1) {
        const u32 m = (l + r) >> 1;
        if (v.ibase[m] <= wi) l = m;
        else r = m;
    }
    it->cloud = cloud;
    it->j = l;
    it->r = wi - v.ibase[l];
    it->idx = l * stride;
    it->lo = v.nlo[it->idx];
    it->hi = v.nlo[it->idx + stride];
    return true;
}

__global__ void __launch_bounds__(256) gb_count(GBArgs a) {
    const u32 cloud = blockIdx.y;
    GBView v = gb_view(a, cloud);
    const u32 wi = blockIdx.x * GB_WPB + warp_id(), lane = lane_id();
    GBItem it;
    if (!gb_item(a, v, cloud, wi, &it)) return;
    const float val = __uint_as_float(v.A0[it.j]);
    const float *col = v.q + (size_t)v.A1[it.j] * a.npad;
    const u32 s0 = it.lo + it.r * GB_WCH, s1 = min(it.hi, s0 + GB_WCH);
    u32 cnt = 0;
    for (u32 i = s0 + lane; i < s1; i += 32) cnt += (col[i] < val) ? 1u : 0u;
    cnt = __reduce_add_sync(FULL, cnt);
    if (lane == 0) v.part[wi] = cnt;
}

__global__ void __launch_bounds__(128) gb_scan(GBArgs a) {
    const u32 nn = 1u << a.lvl;
    const u32 gw = blockIdx.x * (blockDim.x >> 5) + warp_id();
    if (gw >= a.B * nn) return;
    const u32 cloud = gw / nn, j = gw - cloud * nn, lane = lane_id();
    GBView v = gb_view(a, cloud);
    const u32 b0 = v.ibase[j], b1 = v.ibase[j + 1];
    u32 run = 0;
    for (u32 e0 = b0; e0 < b1; e0 += 32) {
        const u32 e = e0 + lane;
        const u32 x = e < b1 ? v.part[e] : 0u;
        u32 inc = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 y = __shfl_up_sync(FULL, inc, o);
            if ((int)lane >= o) inc += y;
        }
        if (e < b1) v.part[e] = run + inc - x;
        run += __shfl_sync(FULL, inc, 31);
    }
    if (lane == 0 && b1 > b0) v.A2[j] = run;
}

__global__ void __launch_bounds__(256) gb_rank(GBArgs a) {
    const u32 cloud = blockIdx.y;
    GBView v = gb_view(a, cloud);
    const u32 wi = blockIdx.x * GB_WPB + warp_id(), lane = lane_id();
    GBItem it;
    if (!gb_item(a, v, cloud, wi, &it)) return;
    const float val = __uint_as_float(v.A0[it.j]);
    const float *col = v.q + (size_t)v.A1[it.j] * a.npad;
    const u32 lo = it.lo, hi = it.hi, m = v.A2[it.j];
    const u32 s0 = lo + it.r * GB_WCH, s1 = min(hi, s0 + GB_WCH);
    u32 base = v.part[wi];
    u32 gl = 0;
    for (u32 i0 = s0; i0 < s1; i0 += 32) {
        const u32 i = i0 + lane;
        const bool in = i < s1;
        const bool f = in && (col[i] < val);
        const u32 mask = __ballot_sync(FULL, f);
        const u32 pre = base + __popc(mask & ((1u << lane) - 1u));
        if (in) {
            if (i < lo + m) {
                if (!f) {
                    v.scr[lo + (i - lo) - pre] = i;
                    ++gl;
                }
            } else if (f) {
                v.scr[hi - m + pre] = i;
            }
        }
        base += __popc(mask);
    }
    gl = __reduce_add_sync(FULL, gl);
    if (lane == 0 && gl) atomicAdd(&v.A3[it.j], gl);
}

__global__ void __launch_bounds__(256) gb_swap(GBArgs a) {
    const u32 cloud = blockIdx.y;
    GBView v = gb_view(a, cloud);
    const u32 wi = blockIdx.x * GB_WPB + warp_id(), lane = lane_id();
    GBItem it;
    if (!gb_item(a, v, cloud, wi, &it)) return;
    const u32 stride = a.S >> a.lvl, half = stride >> 1;
    const u32 lo = it.lo, hi = it.hi, count = hi - lo;
    const u32 g = v.A3[it.j], m = v.A2[it.j];
    const u32 k1 = min(g, (it.r + 1) * GB_WCH);
    // four pairs per lane in flight: the swap is a chain of dependent random accesses (list -> rows) into L2 / HBM
    for (u32 kk0 = it.r * GB_WCH + lane; kk0 < k1; kk0 += 128) {
        u32 pa[4], pb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const u32 kk = kk0 + 32 * u;
            pa[u] = kk < k1 ? v.scr[lo + kk] : 0xffffffffu;
            pb[u] = kk < k1 ? v.scr[hi - 1 - kk] : 0xffffffffu;
        }
        for (u32 c = 0; c < a.dim; ++c) {
            float *col = v.q + (size_t)c * a.npad;
            float xa[4], xb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (pa[u] != 0xffffffffu) xa[u] = col[pa[u]], xb[u] = col[pb[u]];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (pa[u] != 0xffffffffu) col[pa[u]] = xb[u], col[pb[u]] = xa[u];
        }
        u32 ia[4], ib[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (pa[u] != 0xffffffffu) ia[u] = v.perm[pa[u]], ib[u] = v.perm[pb[u]];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (pa[u] != 0xffffffffu) v.perm[pa[u]] = ib[u], v.perm[pb[u]] = ia[u];
    }
    if (it.r == 0) {
        const u32 lim = m == 0 ? 1u : (m == count ? count - 1 : m);   // KDTreeBase.h:142-146
        if (lane == 0) v.nlo[it.idx + half] = lo + lim;
        if (lane < 2 * a.dim) {
            const int init = (lane < a.dim) ? 0x7fffffff : (int)0x80000000;
            v.box[(size_t)it.idx * 2 * a.dim + lane] = init;
            v.box[(size_t)(it.idx + half) * 2 * a.dim + lane] = init;
        }
    }
}

template <int DIM>
__global__ void __launch_bounds__(256) gb_box(GBArgs a) {
    const u32 cloud = blockIdx.y;
    GBView v = gb_view(a, cloud);
    const u32 wi = blockIdx.x * GB_WPB + warp_id();
    GBItem it;
    if (!gb_item(a, v, cloud, wi, &it)) return;
    const u32 half = (a.S >> a.lvl) >> 1;
    const u32 sp = v.nlo[it.idx + half];
    const u32 s0 = it.lo + it.r * GB_WCH, s1 = min(it.hi, s0 + GB_WCH);
    box_range<DIM>(v.q, a.npad, a.dim, s0, s1, sp, v.box + (size_t)it.idx * 2 * a.dim,
                   v.box + (size_t)(it.idx + half) * 2 * a.dim);
}

// boxes: ordered ints -> floats, in place
__global__ void __launch_bounds__(256) gb_finish(GBArgs a) {
    const u32 cloud = blockIdx.y;
    GBView v = gb_view(a, cloud);
    const u32 total = a.S * 2 * a.dim;
    for (u32 e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x)
        reinterpret_cast<float *>(v.box)[e] = ord2f(v.box[e]);
}

// ======================================================================================================
//  host side
// ======================================================================================================
static int pad_dim_g(int dim) { return dim <= 2 ? 2 : dim == 3 ? 3 : dim == 4 ? 4 : dim <= 6 ? 6 : 8; }

static u32 gb_wmax(size_t n, size_t S) {
    size_t w = n / GB_WCH + S + 1;
    return (u32)((w + GB_WPB - 1) / GB_WPB * GB_WPB);
}

size_t kd_gridbuild_aux_bytes(size_t n, size_t dim, size_t h) {
    (void)dim;
    const size_t S = (size_t)1 << h;
    size_t b = (6 * S + 32 + gb_wmax(n, S)) * 4;
    return (b + 255) & ~(size_t)255;
}

template <int DIM>
static void gb_launch_dim(const GBArgs &a, dim3 gi, bool root, cudaStream_t st) {
    if (root) gb_rootbox<DIM><<<gi, 256, 0, st>>>(a);
    else gb_box<DIM><<<gi, 256, 0, st>>>(a);
}

static void gb_box_dispatch(const GBArgs &a, dim3 gi, bool root, cudaStream_t st) {
    switch (pad_dim_g((int)a.dim)) {
        case 2: gb_launch_dim<2>(a, gi, root, st); break;
        case 3: gb_launch_dim<3>(a, gi, root, st); break;
        case 4: gb_launch_dim<4>(a, gi, root, st); break;
        case 6: gb_launch_dim<6>(a, gi, root, st); break;
        default: gb_launch_dim<8>(a, gi, root, st); break;
    }
}

cudaError_t launch_kd_gridbuild(const float *pts, unsigned char *region, size_t region_stride, unsigned char *aux,
                                u32 B, u32 n, u32 dim, u32 h, cudaStream_t st) {
    GBArgs a;
    a.pts = pts;
    a.region = region;
    a.region_stride = region_stride;
    a.aux = aux;
    a.aux_stride = kd_gridbuild_aux_bytes(n, dim, h);
    a.B = B;
    a.n = n;
    a.npad = (n + 31) & ~31u;
    a.dim = dim;
    a.h = h;
    a.S = 1u << h;
    a.nlo_pad = (a.S + 1 + 31) & ~31u;
    a.wmax = gb_wmax(n, a.S);
    a.lvl = 0;
    const u32 stage_blocks = (u32)std::min<size_t>(((size_t)n * dim + 1023) / 1024, 1024);
    gb_stage<<<dim3(stage_blocks, B), 256, 0, st>>>(a);
    count_launch();
    gb_box_dispatch(a, dim3((n + GB_WCH * GB_WPB - 1) / (GB_WCH * GB_WPB), B), true, st);
    count_launch();
    const dim3 gi(a.wmax / GB_WPB, B);
    for (u32 lvl = 0; lvl < h; ++lvl) {
        a.lvl = lvl;
        const u32 nodes = B << lvl;
        gb_split<<<(nodes + 3) / 4, 128, 0, st>>>(a);
        gb_items<<<B, 32, 0, st>>>(a);
        gb_count<<<gi, 256, 0, st>>>(a);
        gb_scan<<<(nodes + 3) / 4, 128, 0, st>>>(a);
        gb_rank<<<gi, 256, 0, st>>>(a);
        gb_swap<<<gi, 256, 0, st>>>(a);
        gb_box_dispatch(a, gi, false, st);
        for (int i = 0; i < 7; ++i) count_launch();
    }
    gb_finish<<<dim3((a.S * 2 * dim + 255) / 256, B), 256, 0, st>>>(a);
    count_launch();
    return cudaGetLastError();
}

}  // namespace fps
