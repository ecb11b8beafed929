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
// The item kernels are memory-latency programs: whatever an item loads is loaded BEFORE its first store (gb_rank: the
// item's column values; gb_swap: every row value of a batch of pairs) -- a store may alias a later load as far as the
// compiler knows, so a load behind a store waits for it: one memory round trip per loop iteration otherwise.
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
    u32 psum, ntmax;   // this level's split values come from the two-phase sum (gb_ptiles / gb_psum_a / gb_psum_b); tiles per cloud at most
};

struct GBView {
    float *q;
    u32 *scr, *perm, *nlo;
    int *box;
    u32 *A0, *A1, *A2, *A3, *nitems, *ibase, *part;
    u32 *ptb;          // two-phase sum: first tile of every node of the level (+ total)
    double *dsum;      // double-precision sum of every 512-element tile (only to GUESS the binade the running sum is in)
    SeqTileRec *recs;  // the tile's integer record under that guess (seqsum.cuh)
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
    unsigned char *px = reinterpret_cast<unsigned char *>(v.part + a.wmax);
    px += (16 - (reinterpret_cast<uintptr_t>(px) & 15)) & 15;
    v.ptb = reinterpret_cast<u32 *>(px);
    px += ((size_t)(a.S + 4) * 4 + 15) & ~(size_t)15;
    v.dsum = reinterpret_cast<double *>(px);
    px += (size_t)a.ntmax * 8;
    px += (16 - (reinterpret_cast<uintptr_t>(px) & 15)) & 15;
    v.recs = reinterpret_cast<SeqTileRec *>(px);
    return v;
}

// ---- stage: row-major -> SoA, identity permutation, slot table, root box reset ---------------------------------
__global__ void __launch_bounds__(256) gb_stage(GBArgs a) {
    // row-major -> SoA through a shared-memory tile of 256 points: coalesced reads of the rows, coalesced writes of every
    // column (a thread-per-float transpose writes 44-byte pieces of three columns per warp: 3.1 ms for 4096 x 100 k x 3)
    constexpr u32 TP = 256;
    __shared__ float tile[TP * 8 + 8];
    const u32 cloud = blockIdx.y, tid = threadIdx.x;
    GBView v = gb_view(a, cloud);
    const float *g = a.pts + (size_t)cloud * a.n * a.dim;
    if (a.dim <= 8) {
        for (u32 i0 = blockIdx.x * TP; i0 < a.n; i0 += gridDim.x * TP) {
            const u32 cnt = min(TP, a.n - i0), nf = cnt * a.dim;
            const float *src = g + (size_t)i0 * a.dim;
            for (u32 f = tid; f < nf; f += 256) tile[f] = src[f];
            __syncthreads();
            if (tid < cnt) {
                for (u32 c = 0; c < a.dim; ++c) v.q[(size_t)c * a.npad + i0 + tid] = tile[tid * a.dim + c];
                v.perm[i0 + tid] = i0 + tid;
            }
            __syncthreads();
        }
    } else {
        const u32 total = a.n * a.dim;
        for (u32 f = blockIdx.x * blockDim.x + tid; f < total; f += gridDim.x * blockDim.x) {
            const u32 i = f / a.dim, c = f - i * a.dim;
            v.q[(size_t)c * a.npad + i] = g[f];
            if (c == 0) v.perm[i] = i;
        }
    }
    if (blockIdx.x == 0) {
        for (u32 s = tid; s <= a.S; s += blockDim.x) v.nlo[s] = (s == a.S) ? a.n : 0u;
        if (tid < 2 * a.dim) v.box[tid] = (tid < a.dim) ? 0x7fffffff : (int)0x80000000;
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

// ---- two-phase sequential sum for LONG columns (one huge cloud: BASELINE.json cfg 4) ---------------------------------------
// The split value is a strictly sequential binary32 sum (KDTreeBase.h:151-158): one dependent FADD per element, 1 M elements
// at the root.  seqsum.cuh turns a 512-element tile into an integer addition when the running sum stays inside one binade,
// but a warp still walks the tiles one after the other (1.6 cycles per element).  Here the expensive part of every tile is
// computed by the WHOLE GRID before the walk, under a guess of the binade the running sum will be in when it gets there:
//   gb_ptiles  one warp per cloud: tiles per node of this level, exclusive prefix
//   gb_psum_a  one warp per tile: the tile's sum in double precision
//   gb_psum_b  one warp per tile: guess = binade of (double prefix of the tiles before it), the tile's integer record
//              under that guess (seq_sum_tile_record: increments for an even / odd incoming sum, prefix bounds)
//   gb_split   one warp per node walks the records: ~12 dependent instructions per tile instead of ~800; a tile whose
//              guess was wrong, or that leaves its binade, or that holds an element too large for the integer model is
//              summed by the plain chain.  The result is the sequential sum bit for bit either way.
constexpr u32 PS_TILE = 512;

__global__ void __launch_bounds__(32) gb_ptiles(GBArgs a) {
    const u32 cloud = blockIdx.x, nn = 1u << a.lvl, stride = a.S >> a.lvl, lane = lane_id();
    GBView v = gb_view(a, cloud);
    u32 run = 0;
    for (u32 j0 = 0; j0 < nn; j0 += 32) {
        const u32 j = j0 + lane;
        u32 x = 0;
        if (j < nn) {
            const u32 cnt = v.nlo[j * stride + stride] - v.nlo[j * stride];
            x = cnt >= 2 ? cnt / PS_TILE : 0u;
        }
        u32 inc = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 y = __shfl_up_sync(FULL, inc, o);
            if ((int)lane >= o) inc += y;
        }
        if (j < nn) v.ptb[j] = run + inc - x;
        run += __shfl_sync(FULL, inc, 31);
    }
    if (lane == 0) v.ptb[nn] = run;
}

// tile wi of the cloud -> (node, tile of the node, first element); false if there is no such tile
__device__ __forceinline__ bool gb_ptile(const GBArgs &a, const GBView &v, u32 wi, u32 *node, u32 *r, const float **src) {
    const u32 nn = 1u << a.lvl, stride = a.S >> a.lvl;
    if (wi >= v.ptb[nn]) return false;
    u32 l = 0, rr = nn;   // last node with ptb[j] <= wi
    while (rr - l > 1) {
        const u32 m = (l + rr) >> 1;
        if (v.ptb[m] <= wi) l = m;
        else rr = m;
    }
    const u32 idx = l * stride;
    const int *b = v.box + (size_t)idx * 2 * a.dim;
    u32 sd = 0;
    float span = 0.0f;
    for (u32 c = 0; c < a.dim; ++c) {   // the split dimension, as gb_split finds it (KDTreeBase.h:160-179)
        const float s = __fsub_rn(ord2f(b[a.dim + c]), ord2f(b[c]));
        if (s > span) {
            span = s;
            sd = c;
        }
    }
    *node = l;
    *r = wi - v.ptb[l];
    *src = v.q + (size_t)sd * a.npad + v.nlo[idx] + (size_t)(wi - v.ptb[l]) * PS_TILE;
    return true;
}

__global__ void __launch_bounds__(256) gb_psum_a(GBArgs a) {
    const u32 cloud = blockIdx.y, lane = lane_id();
    GBView v = gb_view(a, cloud);
    const u32 wi = blockIdx.x * 8 + warp_id();
    u32 node, r;
    const float *src;
    if (!gb_ptile(a, v, wi, &node, &r, &src)) return;
    double acc = 0.0;
#pragma unroll
    for (u32 k = 0; k < PS_TILE / 32; ++k) acc += (double)src[k * 32 + lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o);
    if (lane == 0) v.dsum[wi] = acc;
}

__global__ void __launch_bounds__(256) gb_psum_b(GBArgs a) {
    __shared__ __align__(16) float tile[8][PS_TILE];
    const u32 cloud = blockIdx.y, lane = lane_id(), warp = warp_id();
    GBView v = gb_view(a, cloud);
    const u32 wi = blockIdx.x * 8 + warp;
    u32 node, r;
    const float *src;
    if (!gb_ptile(a, v, wi, &node, &r, &src)) return;
#pragma unroll
    for (u32 k = 0; k < PS_TILE / 32; ++k) tile[warp][k * 32 + lane] = src[k * 32 + lane];
    // the sum of everything before this tile, well enough to know its binade
    double pre = 0.0;
    const double *ds = v.dsum + v.ptb[node];
    for (u32 t = lane; t < r; t += 32) pre += ds[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(FULL, pre, o);
    const u32 ef = (__float_as_uint((float)pre) >> 23) & 0xffu;
    __syncwarp();
    int t0 = 0, t1 = 0, lo = 0, hi = 0;
    const bool ok = seq_sum_tile_record<PS_TILE / 32>(smem_u32(tile[warp]), ef, t0, t1, lo, hi);
    if (lane == 0) {
        SeqTileRec rec;
        rec.ef = ok ? ef : 0u;
        rec.tot0 = t0, rec.tot1 = t1, rec.lo = lo, rec.hi = hi;
        rec.pad[0] = rec.pad[1] = rec.pad[2] = 0u;
        v.recs[wi] = rec;
    }
}

// the walk over the records of one node (one warp; `ring`: 512 floats of warp-private shared memory for the chain fall-back).
// While consecutive records agree on the binade the running sum is kept as the INTEGER k (sum = k * ulp): a tile is then a
// parity select, one integer add and the bound check -- no float <-> int conversion on the chain.
__device__ __forceinline__ float seq_sum_records(const float *src, u32 count, const SeqTileRec *recs, float *ring) {
    const u32 lane = lane_id(), ntile = count / PS_TILE;
    float sum = 0.0f;
    bool fast = false;   // k / cef hold the running sum
    u32 cef = 0;
    int k = 0;
    auto to_float = [&]() {
        if (fast) sum = __fmul_rn(__int2float_rn(k), __uint_as_float((cef - 23u) << 23));   // exact
        fast = false;
    };
    for (u32 t0 = 0; t0 < ntile; t0 += 32) {
        SeqTileRec r;
        r.ef = 0u, r.tot0 = r.tot1 = r.lo = r.hi = 0;
        if (t0 + lane < ntile) r = recs[t0 + lane];
        const u32 nb = min(32u, ntile - t0);
        for (u32 j = 0; j < nb; ++j) {
            const u32 ef = __shfl_sync(FULL, r.ef, j);
            const int a0 = __shfl_sync(FULL, r.tot0, j), a1 = __shfl_sync(FULL, r.tot1, j);
            const int lo = __shfl_sync(FULL, r.lo, j), hi = __shfl_sync(FULL, r.hi, j);
            if (!fast || ef != cef) {   // (re-)enter the integer form in this record's binade, if the sum really is in it
                to_float();
                if (ef != 0u && ((__float_as_uint(sum) >> 23) & 0xffu) == ef) {
                    k = __float2int_rn(__fmul_rn(sum, __uint_as_float((277u - ef) << 23)));   // exact, 2^23 <= |k| < 2^24
                    cef = ef;
                    fast = true;
                }
            }
            bool ok = fast;
            if (ok) {
                const int klo = k + lo, khi = k + hi;
                ok = k > 0 ? (klo > (1 << 23) && khi < (1 << 24)) : (khi < -(1 << 23) && klo > -(1 << 24));
            }
            if (ok) {
                k += (k & 1) ? a1 : a0;   // (the bounds include the tile's last prefix: k stays inside the binade)
                continue;
            }
            to_float();
            const float *ts = src + (size_t)(t0 + j) * PS_TILE;   // the plain chain over this tile
#pragma unroll
            for (u32 kk = 0; kk < PS_TILE / 32; ++kk) ring[kk * 32 + lane] = ts[kk * 32 + lane];
            __syncwarp();
            sum = sq_chain16(smem_u32(ring), PS_TILE, sum);
            __syncwarp();
        }
    }
    to_float();
    // tail (< 512 values): lanes fetch 32 at a time, the chain consumes them through shuffles
    for (u32 i = ntile * PS_TILE; i < count; i += 32) {
        const float x = (i + lane < count) ? src[i + lane] : 0.0f;
        const u32 m = min(32u, count - i);
        for (u32 j = 0; j < m; ++j) sum = __fadd_rn(sum, __shfl_sync(FULL, x, j));
    }
    return sum;
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
    const float sum = a.psum ? seq_sum_records(v.q + (size_t)sd * a.npad + lo, count, v.recs + v.ptb[j], ring[warp_id()])
                             : seq_sum_tma(v.q + (size_t)sd * a.npad + lo, count, ring[warp_id()], bars[warp_id()], phase);
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
    // the item's column values first, eight loads in flight per lane: the list stores below may alias the column as far as
    // the compiler knows, so a load inside the loop would wait for the stores before it -- one memory round trip per 32
    // positions (measured on 4096 clouds x 100 k points: 1.08 ms per level; the column is 1.6 GB = 0.25 ms of HBM time)
    constexpr u32 RB = 8;
    for (u32 j0 = s0; j0 < s1; j0 += 32 * RB) {
        float cv[RB];
#pragma unroll
        for (u32 u = 0; u < RB; ++u) {
            const u32 i = j0 + 32 * u + lane;
            cv[u] = i < s1 ? __ldg(col + i) : 0.0f;
        }
#pragma unroll
        for (u32 u = 0; u < RB; ++u) {
            const u32 i = j0 + 32 * u + lane;
            if (j0 + 32 * u >= s1) break;   // (uniform)
            const bool in = i < s1;
            const bool f = in && (cv[u] < val);
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
    }
    gl = __reduce_add_sync(FULL, gl);
    if (lane == 0 && gl) atomicAdd(&v.A3[it.j], gl);
}

template <int DIM>
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
    // The swap is a chain of dependent random accesses (list -> rows) into L2 / HBM.  Two pairs per lane, and EVERY row value
    // of the batch (all components + the permutation) is loaded before the first one is stored: the stores may alias the
    // loads as far as the compiler knows, so a load behind a store waits for it -- one memory round trip per component
    // otherwise
    constexpr int SP = 2;   // (pairs per lane in flight; 4096 clouds x 100 k x 3, build ms: 1 -> 45.3, 2 -> 39.2, 4 -> 40.8, 8 -> 57.8: registers)
    for (u32 kk0 = it.r * GB_WCH + lane; kk0 < k1; kk0 += 32 * SP) {
        u32 pa[SP], pb[SP];
#pragma unroll
        for (int u = 0; u < SP; ++u) {
            const u32 kk = kk0 + 32 * u;
            pa[u] = kk < k1 ? v.scr[lo + kk] : 0xffffffffu;
            pb[u] = kk < k1 ? v.scr[hi - 1 - kk] : 0xffffffffu;
        }
        float xa[DIM][SP], xb[DIM][SP];
        u32 ia[SP], ib[SP];
#pragma unroll
        for (int c = 0; c < DIM; ++c)
            if (c < (int)a.dim) {
                const float *col = v.q + (size_t)c * a.npad;
#pragma unroll
                for (int u = 0; u < SP; ++u)
                    if (pa[u] != 0xffffffffu) xa[c][u] = col[pa[u]], xb[c][u] = col[pb[u]];
            }
#pragma unroll
        for (int u = 0; u < SP; ++u)
            if (pa[u] != 0xffffffffu) ia[u] = v.perm[pa[u]], ib[u] = v.perm[pb[u]];
#pragma unroll
        for (int c = 0; c < DIM; ++c)
            if (c < (int)a.dim) {
                float *col = v.q + (size_t)c * a.npad;
#pragma unroll
                for (int u = 0; u < SP; ++u)
                    if (pa[u] != 0xffffffffu) col[pa[u]] = xb[c][u], col[pb[u]] = xa[c][u];
            }
#pragma unroll
        for (int u = 0; u < SP; ++u)
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

static u32 gb_ntmax(size_t n) { return (u32)(n / PS_TILE + 1); }

size_t kd_gridbuild_aux_bytes(size_t n, size_t dim, size_t h) {
    (void)dim;
    const size_t S = (size_t)1 << h;
    size_t b = (6 * S + 32 + gb_wmax(n, S)) * 4;
    b += 16 + (((S + 4) * 4 + 15) & ~(size_t)15) + (size_t)gb_ntmax(n) * 8 + 16 + (size_t)gb_ntmax(n) * sizeof(SeqTileRec);   // two-phase sum
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
    a.ntmax = gb_ntmax(n);
    a.psum = 0;
    a.lvl = 0;
    const u32 stage_blocks = (u32)std::min<size_t>(((size_t)n + 255) / 256, 1024);
    gb_stage<<<dim3(stage_blocks, B), 256, 0, st>>>(a);
    count_launch();
    gb_box_dispatch(a, dim3((n + GB_WCH * GB_WPB - 1) / (GB_WCH * GB_WPB), B), true, st);
    count_launch();
    const dim3 gi(a.wmax / GB_WPB, B);
    for (u32 lvl = 0; lvl < h; ++lvl) {
        a.lvl = lvl;
        const u32 nodes = B << lvl;
        // few long columns (the top levels of one huge cloud): the whole grid prepares the tiles of the sequential sum first
        a.psum = (tuning().psum != 0 && (n >> lvl) >= 16384 && (size_t)nodes <= 4 * 148) ? 1u : 0u;
        if (a.psum) {
            const dim3 gt((a.ntmax + 7) / 8, B);
            gb_ptiles<<<B, 32, 0, st>>>(a);
            gb_psum_a<<<gt, 256, 0, st>>>(a);
            gb_psum_b<<<gt, 256, 0, st>>>(a);
            for (int i = 0; i < 3; ++i) count_launch();
        }
        gb_split<<<(nodes + 3) / 4, 128, 0, st>>>(a);
        gb_items<<<B, 32, 0, st>>>(a);
        gb_count<<<gi, 256, 0, st>>>(a);
        gb_scan<<<(nodes + 3) / 4, 128, 0, st>>>(a);
        gb_rank<<<gi, 256, 0, st>>>(a);
        if (dim <= 3) gb_swap<3><<<gi, 256, 0, st>>>(a);
        else if (dim <= 6) gb_swap<6><<<gi, 256, 0, st>>>(a);
        else gb_swap<8><<<gi, 256, 0, st>>>(a);
        gb_box_dispatch(a, gi, false, st);
        for (int i = 0; i < 7; ++i) count_launch();
    }
    gb_finish<<<dim3((a.S * 2 * dim + 255) / 256, B), 256, 0, st>>>(a);
    count_launch();
    return cudaGetLastError();
}

}  // namespace fps
