// kdsmall.cu -- kd-line BUILD for batches of small clouds (a cloud and all its bookkeeping in one CTA's shared memory,
// three CTAs per SM at 4096 x 3): the producer of the per-cloud regions the one-warp-per-cloud sampler reads.
//
// What is reproduced bit for bit (reference src/_ext/KDTreeBase.h:84-207, src/_ext/KDLineTree.h:37-39,
// 87-92): the permutation the reference's recursive build leaves in its point array -- split dim = first dim of
// strictly largest box span (KDTreeBase.h:160-179), split value = SEQUENTIAL binary32 sum / count (:151-158), in-place
// Hoare partition in closed form (:123-149: the k-th misplaced position from the left swaps with the k-th from the right,
// degenerate splits clamp to 1 / count-1 without moving anything), tight child boxes (:112-116, 181-207), leaves at depth
// h or at one point.
//
// Against the general kernel (kdline.cu) this one
//   * addresses shared memory only (every pointer is derived from the dynamic shared array: LDS / STS / ATOMS, no
//     generic loads), keeps the permutation on chip as 16-bit indices next to the points,
//   * needs four phases per level instead of five: the child boxes come out of the counting pass (a child is the set of
//     points on its side of the split value, wherever the partition puts them); only a degenerate split, which keeps
//     positions, recomputes its boxes by position,
//   * drops the block barrier below the level that has one node per warp: from there on a warp owns its subtree and
//     __syncwarp is all the ordering it needs,
//   * stages the cloud with 128-bit loads, four in flight per thread,
//   * evaluates the strictly sequential mean tile by tile as an integer prefix scan where the running sum allows it
//     (seqsum.cuh), the dependent FADD chain elsewhere.
#include <cfloat>

#include "common.cuh"
#include "engine.h"
#include "kdcommon.cuh"
#include "seqsum.cuh"

namespace fps {

struct KdSmallArgs {
    const float *pts;        // [B][n][dim]
    unsigned char *region;   // per cloud: [q dim*npad f32][dis npad f32][perm npad u32][nlo pad32(S+1) u32][fbox S*2*dim f32]
    size_t region_stride;
    unsigned short *idx_ws;  // IDXG: per resident CTA, permutation + misplaced-position scratch (2 * npad u16) in global memory
    u32 B, n, dim, h;
};

// tight box of positions [s0, s1) by one warp, written (not folded) to box[0..2*dim): lows then highs, ordered ints
template <int DIM>
__device__ __forceinline__ void ks_box_write(const float *q, u32 npad, u32 dim, u32 s0, u32 s1, int *box) {
    const u32 lane = lane_id();
    float mn[DIM], mx[DIM];
#pragma unroll
    for (int c = 0; c < DIM; ++c) mn[c] = __int_as_float(0x7f800000), mx[c] = __int_as_float(0xff800000);
    for (u32 i = s0 + lane; i < s1; i += 32) {
#pragma unroll
        for (int c = 0; c < DIM; ++c)
            if (c < (int)dim) {
                const float v = q[c * npad + i];
                mn[c] = fminf(mn[c], v), mx[c] = fmaxf(mx[c], v);
            }
    }
#pragma unroll
    for (int c = 0; c < DIM; ++c)
        if (c < (int)dim) {
            const int a = __reduce_min_sync(FULL, f2ord(mn[c])), b = __reduce_max_sync(FULL, f2ord(mx[c]));
            if (lane == 0) box[c] = a, box[dim + c] = b;
        }
}

// T threads per CTA: 256 with three CTAs per SM for clouds of a few thousand points, 1024 with one CTA per SM for clouds
// whose coordinates alone fill the SM's shared memory (16 384 x 3: BASELINE.json cfg 3); IDXG moves the two 16-bit index
// arrays of such a cloud to global memory (L2-resident, touched only by the scatter / swap phases).
template <int DIM, int T, bool IDXG>
__global__ void __launch_bounds__(T, T == 256 ? 3 : 1) kdsmall_kernel(KdSmallArgs a, u32 *work_counter) {
    constexpr u32 KS_T = T, KS_NW = T / 32;
    extern __shared__ __align__(16) unsigned char ks_smem[];
    __shared__ u32 cloud_s;
    const u32 tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    const u32 n = a.n, dim = a.dim, h = a.h, S = 1u << h, npad = roundup32(n);
    // per-node scratch is indexed by the node's heap number 2^level + j: below the block-synchronous levels the warps
    // run their subtrees at their own pace, so two warps can be at different levels at the same time
    const u32 PN = S > KS_NW ? S : KS_NW;                // (node, rank) pairs: j * ts + rank < NW on the block-synchronous levels

    float *q = reinterpret_cast<float *>(ks_smem);                       // [dim][npad] SoA, permuted in place
    unsigned short *pm, *scr;   // [npad] position -> original id; [npad] misplaced positions of the level
    u32 *nlo;                   // [S + 1] slot boundaries
    if constexpr (IDXG) {
        pm = a.idx_ws + (size_t)blockIdx.x * 2 * npad;
        scr = pm + npad;
        nlo = reinterpret_cast<u32 *>(q + (size_t)dim * npad);
    } else {
        pm = reinterpret_cast<unsigned short *>(q + (size_t)dim * npad);
        scr = pm + npad;
        nlo = reinterpret_cast<u32 *>(scr + npad);
    }
    int *box = reinterpret_cast<int *>(nlo + S + 1);                     // [S][2][dim] ordered ints
    u32 *nval = reinterpret_cast<u32 *>(box + (size_t)S * 2 * dim);      // [S] split value bits, by heap number
    u32 *nsd = nval + S;                                                 // [S] split dim
    u32 *nm = nsd + S;                                                   // [S] count of '< value'
    u32 *part = nm + S;                                                  // [PN] '<' counts per (node, rank)
    u32 *gpart = part + PN;                                              // [PN] misplaced pairs per (node, rank)

    for (;;) {
        __syncthreads();
        if (tid == 0) cloud_s = atomicAdd(work_counter, 1u);
        __syncthreads();
        const u32 cloud = cloud_s;
        if (cloud >= a.B) break;
        const float *gcloud = a.pts + (size_t)cloud * n * dim;

        // ---- stage: row-major -> SoA with 128-bit loads (four in flight per thread), identity permutation -----------
        {
            const u32 nf = n * dim;
            const u32 nf4 = ((reinterpret_cast<uintptr_t>(gcloud) & 15u) == 0) ? (nf >> 2) : 0u;
            const float4 *g4 = reinterpret_cast<const float4 *>(gcloud);
            for (u32 f0 = tid; f0 < nf4; f0 += 4 * KS_T) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const u32 f4 = f0 + u * KS_T;
                    v[u] = (f4 < nf4) ? __ldg(g4 + f4) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const u32 f4 = f0 + u * KS_T;
                    if (f4 < nf4) {
                        const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const u32 f = 4 * f4 + t;
                            const u32 i = (dim == (u32)DIM) ? f / (u32)DIM : f / dim, c = f - i * dim;
                            q[c * npad + i] = e[t];
                        }
                    }
                }
            }
            for (u32 f = 4 * nf4 + tid; f < nf; f += KS_T) {
                const u32 i = f / dim, c = f - i * dim;
                q[c * npad + i] = gcloud[f];
            }
        }
        for (u32 i = tid; i < n; i += KS_T) pm[i] = (unsigned short)i;
        for (u32 s = tid; s <= S; s += KS_T) nlo[s] = (s == S) ? n : 0u;
        if (tid < 2 * dim) box[tid] = (tid < dim) ? 0x7fffffff : (int)0x80000000;
        __syncthreads();
        {   // root box: every warp folds a sub-range
            const u32 chunk = roundup32((n + KS_NW - 1) / KS_NW);
            const u32 s0 = min(n, warp * chunk), s1 = min(n, s0 + chunk);
            box_fold<DIM>(q, npad, dim, s0, s1, box);
        }
        __syncthreads();

        for (u32 lvl = 0; lvl < h; ++lvl) {
            const u32 nn = 1u << lvl, stride = S >> lvl, half = stride >> 1;
            const bool wl = nn >= KS_NW;                     // warp-local levels: a warp owns whole nodes (its subtree)
            const u32 ts = wl ? 1u : KS_NW / nn;             // warps per node
            const u32 npw = wl ? nn / KS_NW : 1u;            // nodes per warp
            const u32 j0 = wl ? warp * npw : warp / ts;
            const u32 rank = wl ? 0u : warp % ts;
            const u32 pbase = wl ? nn : 0u;                  // part / gpart index: heap number on the warp-local levels

            // ---- P1: split dim + sequential mean (KDTreeBase.h:160-179, 151-158); child boxes reset --------------------
            if (rank == 0) {
                for (u32 t = 0; t < npw; ++t) {
                    const u32 j = j0 + t, idx = j * stride;
                    const u32 lo = nlo[idx], hi = nlo[idx + stride], count = hi - lo;
                    if (count < 2) {   // a leaf already (KDLineTree.h:37-39): it stays in the left slot with its box
                        if (lane == 0) nlo[idx + half] = hi;
                        continue;
                    }
                    int *b = box + (size_t)idx * 2 * dim;
                    u32 sd = 0;
                    float span = 0.0f;
                    for (u32 c = 0; c < dim; ++c) {
                        const float s = __fsub_rn(ord2f(b[dim + c]), ord2f(b[c]));
                        if (s > span) span = s, sd = c;
                    }
                    const float sum = seq_sum_shared<16>(smem_u32(q + sd * npad + lo), count);   // tiles of 512 (seqsum.cuh)
                    const float val = __fdiv_rn(sum, __uint2float_rn(count));
                    __syncwarp();   // every lane has read the node's box
                    if (lane == 0) nval[nn + j] = __float_as_uint(val), nsd[nn + j] = sd;
                    if (lane < 2 * dim) {
                        const int init = (lane < dim) ? 0x7fffffff : (int)0x80000000;
                        b[lane] = init;
                        box[(size_t)(idx + half) * 2 * dim + lane] = init;
                    }
                }
            }
            if (wl) __syncwarp(); else __syncthreads();

            // ---- P2: count '< value' per (node, rank) sub-range and fold the points into the child boxes by side ------
            for (u32 t = 0; t < npw; ++t) {
                const u32 j = j0 + t, idx = j * stride;
                const u32 lo = nlo[idx], hi = nlo[idx + stride], count = hi - lo;
                if (count < 2) continue;
                const float val = __uint_as_float(nval[nn + j]);
                const float *col = q + nsd[nn + j] * npad;
                const u32 chunk = roundup32((count + ts - 1) / ts);
                const u32 s0 = min(hi, lo + rank * chunk), s1 = min(hi, s0 + chunk);
                float mnL[DIM], mxL[DIM], mnR[DIM], mxR[DIM];
                const float PINF = __int_as_float(0x7f800000), NINF = __int_as_float(0xff800000);
#pragma unroll
                for (int c = 0; c < DIM; ++c) mnL[c] = mnR[c] = PINF, mxL[c] = mxR[c] = NINF;
                u32 cnt = 0;
#pragma unroll 4
                for (u32 i = s0 + lane; i < s1; i += 32) {
                    const bool f = col[i] < val;
                    cnt += f ? 1u : 0u;
#pragma unroll
                    for (int c = 0; c < DIM; ++c)
                        if (c < (int)dim) {
                            const float v = q[c * npad + i];
                            if (f) mnL[c] = fminf(mnL[c], v), mxL[c] = fmaxf(mxL[c], v);   // predicated, no selects
                            else mnR[c] = fminf(mnR[c], v), mxR[c] = fmaxf(mxR[c], v);
                        }
                }
                cnt = __reduce_add_sync(FULL, cnt);
                if (lane == 0) part[pbase + j * ts + rank] = cnt;
                if (s0 < s1) {
                    int *bL = box + (size_t)idx * 2 * dim, *bR = box + (size_t)(idx + half) * 2 * dim;
#pragma unroll
                    for (int c = 0; c < DIM; ++c)
                        if (c < (int)dim) {
                            const int aL = __reduce_min_sync(FULL, f2ord(mnL[c])), zL = __reduce_max_sync(FULL, f2ord(mxL[c]));
                            const int aR = __reduce_min_sync(FULL, f2ord(mnR[c])), zR = __reduce_max_sync(FULL, f2ord(mxR[c]));
                            if (lane == 0) {
                                if (ts == 1) {   // the only writer
                                    bL[c] = aL, bL[dim + c] = zL, bR[c] = aR, bR[dim + c] = zR;
                                } else {
                                    atomicMin(bL + c, aL), atomicMax(bL + dim + c, zL);
                                    atomicMin(bR + c, aR), atomicMax(bR + dim + c, zR);
                                }
                            }
                        }
                }
            }
            if (wl) __syncwarp(); else __syncthreads();

            // ---- P3: rank the misplaced positions (KDTreeBase.h:123-149 in closed form) ------------------------------------
            for (u32 t = 0; t < npw; ++t) {
                const u32 j = j0 + t, idx = j * stride;
                const u32 lo = nlo[idx], hi = nlo[idx + stride], count = hi - lo;
                if (count < 2) continue;
                const float val = __uint_as_float(nval[nn + j]);
                const float *col = q + nsd[nn + j] * npad;
                const u32 pv = (lane < ts) ? part[pbase + j * ts + lane] : 0u;
                const u32 m = __reduce_add_sync(FULL, pv);
                u32 base = __reduce_add_sync(FULL, lane < rank ? pv : 0u);
                if (rank == 0 && lane == 0) nm[nn + j] = m;
                u32 gl = 0;
                if (m != 0 && m != count) {
                    const u32 chunk = roundup32((count + ts - 1) / ts);
                    const u32 s0 = min(hi, lo + rank * chunk), s1 = min(hi, s0 + chunk);
#pragma unroll 4
                    for (u32 i0 = s0; i0 < s1; i0 += 32) {
                        const u32 i = i0 + lane;
                        const bool in = i < s1;
                        const bool f = in && (col[in ? i : s0] < val);
                        const u32 mask = __ballot_sync(FULL, f);
                        const u32 pre = base + __popc(mask & ((1u << lane) - 1u));
                        if (in) {
                            if (i < lo + m) {
                                if (!f) scr[i - pre] = (unsigned short)i, ++gl;   // k-th '>=' from the left: lo + k
                            } else if (f) {
                                scr[hi - m + pre] = (unsigned short)i;           // '<' on the right, ascending towards hi
                            }
                        }
                        base += __popc(mask);
                    }
                    gl = __reduce_add_sync(FULL, gl);
                }
                if (lane == 0) gpart[pbase + j * ts + rank] = gl;
            }
            if (wl) __syncwarp(); else __syncthreads();

            // ---- P4: swaps (k-th misplaced from the left with k-th from the right), child boundary ---------------------
            for (u32 t = 0; t < npw; ++t) {
                const u32 j = j0 + t, idx = j * stride;
                const u32 lo = nlo[idx], hi = nlo[idx + stride], count = hi - lo;
                if (count < 2) continue;
                const u32 m = nm[nn + j];
                const u32 g = __reduce_add_sync(FULL, (lane < ts) ? gpart[pbase + j * ts + lane] : 0u);
                for (u32 kk = rank * 32 + lane; kk < g; kk += ts * 32) {
                    const u32 pa = scr[lo + kk], pb = scr[hi - 1 - kk];
#pragma unroll
                    for (int c = 0; c < DIM; ++c)
                        if (c < (int)dim) {
                            float *cc = q + c * npad;
                            const float xa = cc[pa], xb = cc[pb];
                            cc[pa] = xb, cc[pb] = xa;
                        }
                    const unsigned short ia = pm[pa], ib = pm[pb];
                    pm[pa] = ib, pm[pb] = ia;
                }
                if (rank == 0) {
                    const u32 lim = m == 0 ? 1u : (m == count ? count - 1 : m);
                    if (lane == 0) nlo[idx + half] = lo + lim;
                    if (m == 0 || m == count) {   // clamped split: nothing moved, the children are position ranges
                        ks_box_write<DIM>(q, npad, dim, lo, lo + lim, box + (size_t)idx * 2 * dim);
                        ks_box_write<DIM>(q, npad, dim, lo + lim, hi, box + (size_t)(idx + half) * 2 * dim);
                    }
                }
            }
            if (wl) __syncwarp(); else __syncthreads();
        }
        __syncthreads();

        // ---- export the permuted cloud, the permutation, slot boundaries and boxes into the cloud's region ---------------
        unsigned char *rg = a.region + (size_t)cloud * a.region_stride;
        float *r_q = reinterpret_cast<float *>(rg);
        u32 *r_perm = reinterpret_cast<u32 *>(rg) + (size_t)(dim + 1) * npad;
        u32 *r_nlo = r_perm + npad;
        float *r_box = reinterpret_cast<float *>(r_nlo + ((S + 1 + 31) & ~31u));
        for (u32 c = 0; c < dim; ++c)
            for (u32 i = tid; i < n; i += KS_T) r_q[(size_t)c * npad + i] = q[c * npad + i];
        for (u32 i = tid; i < n; i += KS_T) r_perm[i] = pm[i];
        for (u32 s = tid; s <= S; s += KS_T) r_nlo[s] = nlo[s];
        for (u32 e = tid; e < S * 2 * dim; e += KS_T) r_box[e] = ord2f(box[e]);
    }
}

// build-only entry (fps_b200_kdline_build_dev): copy permutation / slot boundaries / boxes out of the regions
__global__ void kdsmall_export_kernel(const unsigned char *region, size_t region_stride, u32 n, u32 dim, u32 h, u32 *perm_out,
                                      u32 *leaf_lo_out, float *leaf_box_out) {
    const u32 cloud = blockIdx.x, S = 1u << h, npad = roundup32(n);
    const unsigned char *rg = region + (size_t)cloud * region_stride;
    const u32 *r_perm = reinterpret_cast<const u32 *>(rg) + (size_t)(dim + 1) * npad;
    const u32 *r_nlo = r_perm + npad;
    const float *r_box = reinterpret_cast<const float *>(r_nlo + ((S + 1 + 31) & ~31u));
    if (perm_out)
        for (u32 i = threadIdx.x; i < n; i += blockDim.x) perm_out[(size_t)cloud * n + i] = r_perm[i];
    if (leaf_lo_out)
        for (u32 s = threadIdx.x; s <= S; s += blockDim.x) leaf_lo_out[(size_t)cloud * (S + 1) + s] = r_nlo[s];
    if (leaf_box_out)
        for (u32 e = threadIdx.x; e < S * 2 * dim; e += blockDim.x) leaf_box_out[(size_t)cloud * S * 2 * dim + e] = r_box[e];
}

cudaError_t launch_kdsmall_export(const unsigned char *region, size_t region_stride, u32 B, u32 n, u32 dim, u32 h, u32 *perm_out,
                                  u32 *leaf_lo_out, float *leaf_box_out, cudaStream_t st) {
    kdsmall_export_kernel<<<B, 256, 0, st>>>(region, region_stride, n, dim, h, perm_out, leaf_lo_out, leaf_box_out);
    count_launch();
    return cudaGetLastError();
}

// ======================================================================================================
//  host side
// ======================================================================================================
static int ks_pad_dim(int dim) { return dim <= 2 ? 2 : dim == 3 ? 3 : dim == 4 ? 4 : dim <= 6 ? 6 : 8; }

static size_t ks_smem_bytes(size_t n, size_t dim, size_t h, bool idxg, size_t nw) {
    const size_t S = (size_t)1 << h, npad = (n + 31) & ~(size_t)31;
    const size_t PN = S > nw ? S : nw;
    return dim * npad * 4 + (idxg ? 0 : 2 * npad * 2) + ((S + 1) + S * 2 * dim + 3 * S + 2 * PN) * 4 + 16;
}

template <int DIM, int T, bool IDXG>
static cudaError_t ks_occupancy(size_t smem, int *occ) {
    auto kern = kdsmall_kernel<DIM, T, IDXG>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, T, smem);
}

template <int DIM>
static cudaError_t ks_occupancy_d(bool big, size_t smem, int *occ) {
    return big ? ks_occupancy<DIM, 1024, true>(smem, occ) : ks_occupancy<DIM, 256, false>(smem, occ);
}

bool plan_kdsmall(size_t n, size_t dim, size_t h, size_t B, int n_sms, KdSmallPlan *pl) {
    if (dim == 0 || dim > 8 || n == 0 || n > 65535 || h == 0 || h > 8 || B == 0) return false;
    if (tuning().kdsmall == 0) return false;
    size_t smem = ks_smem_bytes(n, dim, h, false, 8);
    bool big = false;
    if (smem > 110 * 1024) {   // fewer than two clouds per SM: one CTA of 1024 threads per SM, index arrays in global memory
        big = true;
        smem = ks_smem_bytes(n, dim, h, true, 32);
        if (smem > 225 * 1024) return false;
    }
    pl->dimp = ks_pad_dim((int)dim);
    pl->smem = smem;
    pl->big = big ? 1 : 0;
    int occ = 0;
    cudaError_t e;
    switch (pl->dimp) {
        case 2: e = ks_occupancy_d<2>(big, smem, &occ); break;
        case 3: e = ks_occupancy_d<3>(big, smem, &occ); break;
        case 4: e = ks_occupancy_d<4>(big, smem, &occ); break;
        case 6: e = ks_occupancy_d<6>(big, smem, &occ); break;
        default: e = ks_occupancy_d<8>(big, smem, &occ); break;
    }
    if (e != cudaSuccess || occ < 1) {
        cudaGetLastError();
        return false;
    }
    size_t grid = (size_t)occ * (size_t)n_sms;
    if (grid > B) grid = B;
    pl->grid = (u32)grid;
    pl->occ = (u32)occ;
    const size_t npad = (n + 31) & ~(size_t)31;
    pl->ws_bytes = 256 + (big ? (size_t)occ * (size_t)n_sms * 2 * npad * sizeof(unsigned short) : 0);
    return true;
}

template <int DIM>
static void ks_launch_d(const KdSmallPlan &pl, const KdSmallArgs &a, u32 *counter, cudaStream_t st) {
    if (pl.big) kdsmall_kernel<DIM, 1024, true><<<pl.grid, 1024, pl.smem, st>>>(a, counter);
    else kdsmall_kernel<DIM, 256, false><<<pl.grid, 256, pl.smem, st>>>(a, counter);
}

// ws: pl.ws_bytes of workspace (256 bytes of scheduler counter, then the index arrays of the big variant)
cudaError_t launch_kdsmall(const KdSmallPlan &pl, const float *pts, unsigned char *region, size_t region_stride,
                           u32 *ws, u32 B, u32 n, u32 dim, u32 h, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(ws, 0, 256, st);
    if (e != cudaSuccess) return e;
    KdSmallArgs a;
    a.pts = pts;
    a.region = region;
    a.region_stride = region_stride;
    a.idx_ws = reinterpret_cast<unsigned short *>(reinterpret_cast<unsigned char *>(ws) + 256);
    a.B = B, a.n = n, a.dim = dim, a.h = h;
    switch (pl.dimp) {
        case 2: ks_launch_d<2>(pl, a, ws, st); break;
        case 3: ks_launch_d<3>(pl, a, ws, st); break;
        case 4: ks_launch_d<4>(pl, a, ws, st); break;
        case 6: ks_launch_d<6>(pl, a, ws, st); break;
        default: ks_launch_d<8>(pl, a, ws, st); break;
    }
    count_launch();
    return cudaGetLastError();
}

}  // namespace fps
