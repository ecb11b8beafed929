// kdline.cu -- QuickFPS kd-line path (bucket_fps_kdline_sampling), one persistent CTA per cloud.
//
// What must be reproduced bit-for-bit is:
//   (1) the permutation the reference's recursive build leaves in its point array
//       (src/_ext/KDTreeBase.h:84-207: first-max-span split dim, SEQUENTIAL binary32 mean as split
//       value, in-place Hoare partition, tight child boxes; leaf rule src/_ext/KDLineTree.h:37-39);
//   (2) exact FPS over that permuted array, started at POSITION start (src/wrapper.hpp:54-55), running
//       distance initialised to FLT_MAX (src/_ext/Point.h:61-65), ties to the LOWEST position
//       (strict '>' at src/_ext/KDNode.h:95-101,153-159 and src/_ext/KDLineTree.h:56-67).
// The reference's bucket bookkeeping (defer / flush lists, KDNode.h:120-166) is an accelerator only;
// here each warp owns buckets, tests them against the new sample with the reference's own
// point-to-box bound (KDNode.h:105-118) and rescans a bucket only when the bound is below the
// bucket's current maximum.  Float rounding is monotone, so a skipped bucket provably has no point
// whose distance would drop: the result equals the eager recurrence exactly.
//
// Build, level by level (node j of level l lives at slot j << (h-l); children reuse slot / slot+half):
//   P1 one warp per node: split dim + sequential f32 sum (lane-broadcast add chain) -> split value
//   P2 teams of warps count '< value' per sub-range          P3 rank misplaced elements (ballot scans)
//   P4 pairwise swaps (k-th misplaced from the left with k-th from the right == the Hoare loop)
//   P5 child boxes via redux.sync min/max on order-preserving ints + shared/global atomics
#include <cfloat>

#include "common.cuh"
#include "engine.h"
#include "kdcommon.cuh"

namespace fps {
__device__ unsigned long long g_kb_dbg[16];
#ifndef KBDBG
#define KBDBG 0   // 1: per-phase clock64 counters of cloud 0 (scripts/time_build.py prints them)
#endif

struct KdFixedSmem {
    u64 wslot[2][32];
    u32 cloud;
};

template <int DIM>
__global__ void __launch_bounds__(1024, 1) kdline_kernel(KdlineArgs a, u32 *work_counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    KdFixedSmem &F = *reinterpret_cast<KdFixedSmem *>(smem_raw);

    const u32 tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    const u32 T = blockDim.x, NW = T >> 5;
    const u32 n = a.n, dim = a.dim, h = a.h, S = 1u << h;
    const u32 npad = roundup32(n);
    const u32 HS = S > 1 ? (S >> 1) : 1;               // max nodes that are ever split in one level
    const u32 PN = (HS > NW ? HS : NW) + 32;           // per-(node,rank) partial counts

    // ---- carve memory: data = q[dim][npad] + scr[npad]; meta = nlo | box | 4 per-node arrays | part ---
    const size_t data_bytes = ((size_t)dim + 1) * npad * 4;
    const size_t meta_bytes = ((size_t)(S + 1) + (size_t)S * 2 * dim + 4 * (size_t)S + PN) * 4;
    float *chainbuf = reinterpret_cast<float *>(smem_raw + ((sizeof(KdFixedSmem) + 15) & ~15)) + (size_t)warp * 256;
    unsigned char *sm = smem_raw + ((sizeof(KdFixedSmem) + 15) & ~15) + (size_t)NW * 1024;
    unsigned char *gw = a.ws + (size_t)blockIdx.x * a.ws_stride;
    u32 *perm_ws = reinterpret_cast<u32 *>(gw);
    gw += (size_t)npad * 4;
    unsigned char *data_base, *meta_base;
    if (a.in_smem & 1) {
        data_base = sm;
        sm += (data_bytes + 15) & ~(size_t)15;
    } else {
        data_base = gw;   // replaced per cloud when a.region is set
        if (!a.region) gw += (data_bytes + 15) & ~(size_t)15;
    }
    meta_base = (a.in_smem & 2) ? sm : gw;
    (void)meta_bytes;
    float *q = reinterpret_cast<float *>(data_base);
    u32 *scr = reinterpret_cast<u32 *>(data_base) + (size_t)dim * npad;
    u32 *nlo = reinterpret_cast<u32 *>(meta_base);
    int *box = reinterpret_cast<int *>(nlo + S + 1);
    u32 *A0 = reinterpret_cast<u32 *>(box + (size_t)S * 2 * dim);  // build: split value ; sample: bucket max dis
    u32 *A1 = A0 + S;                                              // build: split dim   ; sample: bucket max pos
    u32 *A2 = A1 + S;                                              // build: m (count '<')
    u32 *A3 = A2 + S;                                              // build: g (misplaced pairs)
    u32 *part = A3 + S;

    for (;;) {
        // ---- dynamic cloud scheduler ------------------------------------------------------------------
        __syncthreads();
        if (tid == 0) F.cloud = atomicAdd(work_counter, 1u);
        __syncthreads();
        const u32 cloud = F.cloud;
        if (cloud >= a.B) break;

        const float *gcloud = a.pts + (size_t)cloud * n * dim;
        u32 *perm = a.perm_out ? a.perm_out + (size_t)cloud * n : perm_ws;
        u32 *r_nlo = nullptr;
        float *r_box = nullptr;
        float *r_q = nullptr;
        u32 *r_perm = nullptr;
        if (a.region) {  // build-only into the per-cloud region: [q dim*npad][dis npad][perm npad][nlo][fbox]
            unsigned char *rg = a.region + (size_t)cloud * a.region_stride;
            u32 *rscr = reinterpret_cast<u32 *>(rg) + (size_t)dim * npad;
            r_nlo = rscr + 2 * (size_t)npad;
            r_box = reinterpret_cast<float *>(r_nlo + ((S + 1 + 31) & ~31u));
            if (a.in_smem & 1) {   // small cloud: build in shared memory, export the permuted cloud at the end
                r_q = reinterpret_cast<float *>(rg);
                r_perm = rscr + npad;
            } else {               // big cloud: build in place in the region (L2)
                q = reinterpret_cast<float *>(rg);
                scr = rscr;
                perm = rscr + npad;
            }
        }

        // ---- stage: row-major -> SoA, identity permutation ---------------------------------------------
        for (u32 f = tid; f < n * dim; f += T) {
            u32 i = f / dim, c = f - i * dim;
            q[(size_t)c * npad + i] = gcloud[f];
        }
        for (u32 i = tid; i < n; i += T) perm[i] = i;
        for (u32 s = tid; s <= S; s += T) nlo[s] = (s == S) ? n : 0u;
        if (tid < 2 * dim) box[tid] = (tid < dim) ? 0x7fffffff : (int)0x80000000;
        __syncthreads();
        {  // root box: every warp takes a sub-range
            u32 chunk = roundup32((n + NW - 1) / NW);
            u32 s0 = min(n, warp * chunk), s1 = min(n, s0 + chunk);
            if (s0 < s1) box_range<DIM>(q, npad, dim, s0, s1, n, box, box);
        }
        __syncthreads();

#if KBDBG
        unsigned long long kd[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        long long kt0 = clock64();
        kd[0] = 0;
#endif
        // ---- build -----------------------------------------------------------------------------------
        for (u32 lvl = 0; lvl < h; ++lvl) {
#if KBDBG
            long long ka = clock64();
#endif
            const u32 nn = 1u << lvl, stride = S >> lvl, half = stride >> 1;
            const u32 ts = nn < NW ? NW / nn : 1u;  // warps per node
            const u32 nteams = NW / ts;
            const u32 team = warp / ts, rank = warp - team * ts;

            // P1: split dim + sequential mean (KDTreeBase.h:160-179, 151-158)
            if (rank == 0 && team < nteams) {
                for (u32 j = team; j < nn; j += nteams) {
                    const u32 idx = j * stride;
                    const u32 lo = nlo[idx], hi = nlo[idx + stride];
                    if (hi - lo < 2) continue;
                    const int *b = box + (size_t)idx * 2 * dim;
                    u32 sd = 0;
                    float span = 0.0f;
                    for (u32 c = 0; c < dim; ++c) {
                        float s = __fsub_rn(ord2f(b[dim + c]), ord2f(b[c]));
                        if (s > span) {
                            span = s;
                            sd = c;
                        }
                    }
                    const float *col = q + (size_t)sd * npad + lo;
                    const float sum = ((a.in_smem & 1) && !(a.region && !(a.in_smem & 1))) ? seq_sum_smem(col, hi - lo)
                                                                                        : seq_sum_staged(col, hi - lo, chainbuf);
                    float val = __fdiv_rn(sum, __uint2float_rn(hi - lo));
                    if (lane == 0) {
                        A0[j] = __float_as_uint(val);
                        A1[j] = sd;
                        A3[j] = 0;
                    }
                }
            }
            __syncthreads();
#if KBDBG
            { long long kb = clock64(); kd[1] += kb - ka; ka = kb; }
#endif
            // P2: count '< val' per (node, rank) sub-range
            if (team < nteams) {
                for (u32 j = team; j < nn; j += nteams) {
                    const u32 idx = j * stride;
                    const u32 lo = nlo[idx], hi = nlo[idx + stride], count = hi - lo;
                    if (count < 2) continue;
                    const float val = __uint_as_float(A0[j]);
                    const float *col = q + (size_t)A1[j] * npad;
                    const u32 chunk = roundup32((count + ts - 1) / ts);
                    const u32 s0 = min(hi, lo + rank * chunk), s1 = min(hi, s0 + chunk);
                    u32 cnt = 0;
                    for (u32 i = s0 + lane; i < s1; i += 32) cnt += (col[i] < val) ? 1u : 0u;
                    cnt = __reduce_add_sync(FULL, cnt);
                    if (lane == 0) part[j * ts + rank] = cnt;
                }
            }
            __syncthreads();
#if KBDBG
            { long long kb = clock64(); kd[2] += kb - ka; ka = kb; }
#endif
            // P3: rank the misplaced elements (KDTreeBase.h:123-149 in closed form)
            if (team < nteams) {
                for (u32 j = team; j < nn; j += nteams) {
                    const u32 idx = j * stride;
                    const u32 lo = nlo[idx], hi = nlo[idx + stride], count = hi - lo;
                    if (count < 2) continue;
                    const float val = __uint_as_float(A0[j]);
                    const float *col = q + (size_t)A1[j] * npad;
                    u32 pv = (lane < ts) ? part[j * ts + lane] : 0u;
                    const u32 m = __reduce_add_sync(FULL, pv);
                    u32 base = __reduce_add_sync(FULL, lane < rank ? pv : 0u);
                    if (rank == 0 && lane == 0) A2[j] = m;
                    const u32 chunk = roundup32((count + ts - 1) / ts);
                    const u32 s0 = min(hi, lo + rank * chunk), s1 = min(hi, s0 + chunk);
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
                                    scr[lo + (i - lo) - pre] = i;
                                    ++gl;
                                }
                            } else if (f) {
                                scr[hi - m + pre] = i;
                            }
                        }
                        base += __popc(mask);
                    }
                    gl = __reduce_add_sync(FULL, gl);
                    if (lane == 0 && gl) atomicAdd(&A3[j], gl);
                }
            }
            __syncthreads();
#if KBDBG
            { long long kb = clock64(); kd[3] += kb - ka; ka = kb; }
#endif
            // P4: swaps, child boundaries, child box init
            if (team < nteams) {
                for (u32 j = team; j < nn; j += nteams) {
                    const u32 idx = j * stride;
                    const u32 lo = nlo[idx], hi = nlo[idx + stride], count = hi - lo;
                    u32 lim = count;  // count 0 or 1: no split, everything stays in the left child
                    if (count >= 2) {
                        const u32 g = A3[j], m = A2[j];
                        for (u32 kk = rank * 32 + lane; kk < g; kk += ts * 32) {
                            const u32 pa = scr[lo + kk], pb = scr[hi - 1 - kk];
                            for (u32 c = 0; c < dim; ++c) {
                                float *col = q + (size_t)c * npad;
                                float xa = col[pa], xb = col[pb];
                                col[pa] = xb;
                                col[pb] = xa;
                            }
                            u32 ia = perm[pa], ib = perm[pb];
                            perm[pa] = ib;
                            perm[pb] = ia;
                        }
                        lim = m == 0 ? 1u : (m == count ? count - 1 : m);
                    }
                    if (rank == 0) {
                        if (lane == 0) nlo[idx + half] = lo + lim;
                        __syncwarp();
                        if (lane < 2 * dim) {
                            const int init = (lane < dim) ? 0x7fffffff : (int)0x80000000;
                            if (count >= 2) box[(size_t)idx * 2 * dim + lane] = init;  // count<2: left child keeps the box
                            box[(size_t)(idx + half) * 2 * dim + lane] = init;
                        }
                    }
                }
            }
            __syncthreads();
#if KBDBG
            { long long kb = clock64(); kd[4] += kb - ka; ka = kb; }
#endif
            // P5: tight child boxes (KDTreeBase.h:112-116, 181-207)
            if (team < nteams) {
                for (u32 j = team; j < nn; j += nteams) {
                    const u32 idx = j * stride;
                    const u32 lo = nlo[idx], hi = nlo[idx + stride], count = hi - lo;
                    if (count < 2) continue;
                    const u32 sp = nlo[idx + half];
                    const u32 chunk = roundup32((count + ts - 1) / ts);
                    const u32 s0 = min(hi, lo + rank * chunk), s1 = min(hi, s0 + chunk);
                    if (s0 < s1)
                        box_range<DIM>(q, npad, dim, s0, s1, sp, box + (size_t)idx * 2 * dim,
                                       box + (size_t)(idx + half) * 2 * dim);
                }
            }
            __syncthreads();
#if KBDBG
            { long long kb = clock64(); kd[5] += kb - ka; ka = kb; }
#endif
        }

#if KBDBG
        kd[6] = clock64() - kt0;
        if (cloud == 0 && tid == 0) { for (int i = 0; i < 8; ++i) g_kb_dbg[i] = kd[i]; }
#endif
        // ---- leaves: decode boxes to floats; optional export ---------------------------------------------
        float *fbox = reinterpret_cast<float *>(box);
        for (u32 e = tid; e < S * 2 * dim; e += T) fbox[e] = ord2f(box[e]);
        __syncthreads();
        if (a.leaf_lo_out)
            for (u32 s = tid; s <= S; s += T) a.leaf_lo_out[(size_t)cloud * (S + 1) + s] = nlo[s];
        if (a.leaf_box_out)
            for (u32 e = tid; e < S * 2 * dim; e += T) a.leaf_box_out[(size_t)cloud * S * 2 * dim + e] = fbox[e];
        if (r_nlo) {
            for (u32 s = tid; s <= S; s += T) r_nlo[s] = nlo[s];
            for (u32 e = tid; e < S * 2 * dim; e += T) r_box[e] = fbox[e];
        }
        if (r_q) {
            for (u32 c = 0; c < dim; ++c)
                for (u32 i = tid; i < n; i += T) r_q[(size_t)c * npad + i] = q[(size_t)c * npad + i];
            for (u32 i = tid; i < n; i += T) r_perm[i] = perm[i];
        }
        if (!a.out || a.region) continue;

        // ---- sample -------------------------------------------------------------------------------------
        float *dis = reinterpret_cast<float *>(scr);
        for (u32 i = tid; i < n; i += T) dis[i] = FLT_MAX;  // Point.h:61-65
        __syncthreads();
        u32 cur = a.starts ? (u32)a.starts[cloud] : 0u;
        float r[DIM];
#pragma unroll
        for (int c = 0; c < DIM; ++c) r[c] = (c < (int)dim) ? q[(size_t)c * npad + cur] : 0.0f;
        u64 *out = a.out + (size_t)cloud * a.k;
        if (tid == 0) out[0] = perm[cur];
        const u32 slots_w = (S > warp) ? (S - warp + NW - 1) / NW : 0u;  // buckets owned by this warp

        for (u32 t = 1; t < a.k; ++t) {
            const u32 par = t & 1;
            const bool force = (t == 1);  // KDNode::init: every leaf scans the first reference
            u64 wbest = 0;
            for (u32 sb = 0; sb < slots_w; sb += 32) {
                const u32 slot = sb + lane;
                const u32 s = slot * NW + warp;
                const bool valid = slot < slots_w;
                u32 blo = 0, bhi = 0;
                if (valid) {
                    blo = nlo[s];
                    bhi = nlo[s + 1];
                }
                const bool nonempty = bhi > blo;
                float md = 0.0f;
                u32 mp = 0;
                bool need = false;
                if (nonempty) {
                    if (force) {
                        need = true;
                    } else {
                        md = __uint_as_float(A0[s]);
                        mp = A1[s];
                        float bl[DIM], bh[DIM];
#pragma unroll
                        for (int c = 0; c < DIM; ++c) {
                            bl[c] = (c < (int)dim) ? fbox[(size_t)s * 2 * dim + c] : 0.0f;
                            bh[c] = (c < (int)dim) ? fbox[(size_t)s * 2 * dim + dim + c] : 0.0f;
                        }
                        need = boxdist<DIM>(r, bl, bh) < md;
                    }
                }
                u32 mask = __ballot_sync(FULL, need);
                while (mask) {
                    const int src = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const u32 lo_ = __shfl_sync(FULL, blo, src), hi_ = __shfl_sync(FULL, bhi, src);
                    u64 best = 0;
                    for (u32 i = lo_ + lane; i < hi_; i += 32) {
                        float p[DIM];
#pragma unroll
                        for (int c = 0; c < DIM; ++c) p[c] = (c < (int)dim) ? q[(size_t)c * npad + i] : 0.0f;
                        const float d = sqdist<DIM>(p, r);
                        const float v = fminf(dis[i], d);  // std::min(dis, d), Point.h:82-86
                        dis[i] = v;
                        const u64 key = make_key(v, ~i);
                        best = key > best ? key : best;
                    }
                    best = warp_max_key(best);
                    if ((int)lane == src) {
                        md = __uint_as_float((u32)(best >> 32));
                        mp = ~(u32)best;
                        A0[s] = (u32)(best >> 32);
                        A1[s] = mp;
                    }
                }
                const u64 key = nonempty ? make_key(md, ~mp) : 0ull;
                wbest = key > wbest ? key : wbest;
            }
            wbest = warp_max_key(wbest);
            if (lane == 0) F.wslot[par][warp] = wbest;
            __syncthreads();
            u64 kk = (lane < NW) ? F.wslot[par][lane] : 0ull;
            kk = warp_max_key(kk);
            cur = ~(u32)kk;  // lowest position among maxima (KDLineTree.h:56-67)
#pragma unroll
            for (int c = 0; c < DIM; ++c) r[c] = (c < (int)dim) ? q[(size_t)c * npad + cur] : 0.0f;
            if (tid == 0) out[t] = perm[cur];
        }
    }
}

// ======================================================================================================
//  host side
// ======================================================================================================
static int pad_dim_k(int dim) { return dim <= 2 ? 2 : dim == 3 ? 3 : dim == 4 ? 4 : dim <= 6 ? 6 : 8; }

static size_t kd_meta_bytes(size_t S, size_t dim, size_t NW) {
    size_t HS = S > 1 ? S / 2 : 1;
    size_t PN = (HS > NW ? HS : NW) + 32;
    return ((S + 1) + S * 2 * dim + 4 * S + PN) * 4;
}

template <int DIM>
static cudaError_t kd_occupancy(const KdlinePlan &pl, int *occ) {
    auto kern = kdline_kernel<DIM>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, (int)pl.threads, pl.smem);
}

// Decide placement (shared memory vs workspace) and the persistent grid for a kd-line batch.
cudaError_t plan_kdline(size_t n, size_t dim, size_t h, size_t B, int n_sms, KdlinePlan *pl) {
    if (dim == 0 || dim > 8 || n == 0 || h == 0 || h > 24 || B == 0) return cudaErrorInvalidValue;
    const size_t S = (size_t)1 << h;
    const size_t npad = (n + 31) & ~(size_t)31;
    u32 threads = n <= 8192 ? 256 : (n <= 65536 ? 512 : 1024);
    while (threads < 1024 && S > (size_t)threads) threads *= 2;  // at most ~32 buckets per warp
    const size_t NW = threads / 32;
    const size_t fixed = ((sizeof(KdFixedSmem) + 15) & ~(size_t)15) + NW * 1024;  // + per-warp chain staging
    const size_t data = (((dim + 1) * npad * 4) + 15) & ~(size_t)15;
    const size_t meta = (kd_meta_bytes(S, dim, NW) + 15) & ~(size_t)15;
    const size_t cap = 200 * 1024;
    u32 mask = 0;
    size_t smem = fixed;
    if (fixed + data + meta <= cap) {
        mask = 3;
        smem += data + meta;
    } else if (fixed + meta <= cap / 2) {
        mask = 2;
        smem += meta;
    }
    size_t ws = npad * 4;
    if (!(mask & 1)) ws += data;
    if (!(mask & 2)) ws += meta;
    ws = (ws + 255) & ~(size_t)255;
    pl->dimp = pad_dim_k((int)dim);
    pl->threads = threads;
    pl->in_smem = mask;
    pl->smem = smem;
    pl->ws_stride = ws;
    int occ = 0;
    cudaError_t e;
    switch (pl->dimp) {
        case 2: e = kd_occupancy<2>(*pl, &occ); break;
        case 3: e = kd_occupancy<3>(*pl, &occ); break;
        case 4: e = kd_occupancy<4>(*pl, &occ); break;
        case 6: e = kd_occupancy<6>(*pl, &occ); break;
        default: e = kd_occupancy<8>(*pl, &occ); break;
    }
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    size_t grid = (size_t)occ * (size_t)n_sms;
    if (grid > B) grid = B;
    pl->grid = (u32)grid;
    pl->ws_bytes = 256 + grid * ws;
    return cudaSuccess;
}

cudaError_t kb_debug_counters(unsigned long long *out16) { return cudaMemcpyFromSymbol(out16, g_kb_dbg, sizeof(unsigned long long) * 16); }

cudaError_t launch_kdline(const KdlinePlan &pl, KdlineArgs a, unsigned char *ws_base, cudaStream_t st) {
    u32 *counter = reinterpret_cast<u32 *>(ws_base);
    cudaError_t e = cudaMemsetAsync(counter, 0, 256, st);
    if (e != cudaSuccess) return e;
    a.ws = ws_base + 256;
    a.ws_stride = pl.ws_stride;
    a.in_smem = pl.in_smem;
    switch (pl.dimp) {
        case 2: kdline_kernel<2><<<pl.grid, pl.threads, pl.smem, st>>>(a, counter); break;
        case 3: kdline_kernel<3><<<pl.grid, pl.threads, pl.smem, st>>>(a, counter); break;
        case 4: kdline_kernel<4><<<pl.grid, pl.threads, pl.smem, st>>>(a, counter); break;
        case 6: kdline_kernel<6><<<pl.grid, pl.threads, pl.smem, st>>>(a, counter); break;
        default: kdline_kernel<8><<<pl.grid, pl.threads, pl.smem, st>>>(a, counter); break;
    }
    count_launch();
    return cudaGetLastError();
}

}  // namespace fps
