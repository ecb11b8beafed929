// kdline_async.cu -- QuickFPS kd-line SAMPLING for clouds that do not fit one SM's shared memory: one
// thread-block cluster per cloud, CTA 0 = coordinator, CTAs 1..C-1 = scan workers, data (permuted SoA coordinates
// + running distances) resident in L2.
//
// Same observable result as the reference's lazy bucket scheme (src/_ext/KDNode.h:120-166,
// src/_ext/KDLineTree.h:56-85), i.e. exact FPS over the permuted array, ties to the lowest position
//  -- but the K-1 dependent picks no longer wait for the leaf rescans:
//
//   * every bucket is EXACT (max distance, its position and coordinates known; pending references provably do
//     not touch the max point -- the reference's "delaypoints" state, KDNode.h:124-134) or INFLIGHT (a rescan job
//     is out at a worker; only an UPPER BOUND U of the bucket's max is known: the second-largest distance of the
//     last scan, or the old max, whichever the triggering reference leaves standing);
//   * the coordinator picks arg-max over {exact max} U {upper bounds}; a pick is final as soon as the winner is an
//     EXACT bucket whose value is strictly above every upper bound (ties wait), so rescans overlap later picks;
//   * one coordinator thread owns one bucket: per pick it applies the reference's own tests -- point-to-box bound
//     (KDNode.h:105-118) against U / max, distance to the bucket's max point against max (KDNode.h:122-123) -- and
//     either drops the reference, appends it to the bucket's pending list, or ships the list as a job;
//   * jobs and results travel through distributed shared memory (st.shared::cluster + release/acquire flags):
//     no global-memory round trip and no cluster barrier on the pick loop.
// scripts/sim_async.py is the CPU model of this protocol (checked bit-exact against the oracle).
#include <cfloat>

#include "common.cuh"
#include "engine.h"

namespace fps {

constexpr u32 A_RING = 32;       // job slots per worker
constexpr u32 A_MAXR = 12;       // max pending references per bucket
constexpr u32 A_MARK = 0xffffffffu;
constexpr u32 A_QUIT = 0x7fffu;

// ---- DSMEM messaging primitives ---------------------------------------------------------------------------
// jobs   (coordinator -> worker): 16-byte st.async chunks that complete_tx on the slot's mbarrier in the
//         worker's shared memory + one relaxed remote arrive.expect_tx: no fence on the coordinator's critical
//         path, and the worker threads sleep in mbarrier.try_wait instead of spinning.
// results (worker -> coordinator): 16-byte st.shared::cluster.v4 chunks, each stamped with the job's sequence
//         number (a 16-byte aligned vector store lands as one unit), polled with plain 128-bit shared loads.
__device__ __forceinline__ void st_async_v4(u32 caddr, u32 x, u32 y, u32 z, u32 w, u32 cbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(caddr),
                 "r"(x), "r"(y), "r"(z), "r"(w), "r"(cbar)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_remote(u32 cbar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.relaxed.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_cta(u32 bar, u32 parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void st_cluster_v4(u32 caddr, u32 x, u32 y, u32 z, u32 w) {
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(caddr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void st_cluster_u32(u32 caddr, u32 v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(caddr), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 lds_v4_volatile(const void *p) {
    uint4 v;
    asm volatile("ld.volatile.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ u32 lds_volatile(const void *p) {
    u32 v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}

template <int DIM>
struct AJob {   // 16-byte chunks: header, then the references' coordinates packed [r][c]
    u32 hdr;    // bucket | first << 15 | nrefs << 16
    u32 lo, hi, seq;
    float c[(A_MAXR * DIM + 3) / 4 * 4];
};

__host__ __device__ constexpr int ares_chunks(int dim) { return 1 + (dim + 2) / 3; }
template <int DIM>
struct ARes {   // chunk 0 {seq, mx, pos, snd}; chunk i {seq, c[3i-3], c[3i-2], c[3i-1]}
    uint4 ch[ares_chunks(DIM)];
};

constexpr int A_NC = 32;         // candidates per coordinator iteration (one per lane of warp 0)
struct ACand {                   // a warp's k-th largest key
    u64 key;
    u32 bucket;
    u32 pad;
};
template <int DIM>
struct ATab {                    // the iteration's candidates in descending key order (written by warp 0)
    u32 L0;                      // leading candidates that are EXACT and above every bound
    u32 bad;                     // bit j: candidate j conflicts with an earlier one (set by the pair checks)
    u32 pad[2];
    u32 pos[A_NC];
    float val[A_NC];
    float snd[A_NC];
    float c[DIM][A_NC];
};

__device__ __forceinline__ void bar_sync_id(u32 id, u32 nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bar_arrive_id(u32 id, u32 nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// coordinator phase counters of the last launch (thread 0 of cluster 0's coordinator): diagnostics only
__device__ u64 g_async_dbg[16];

struct AsyncArgs {
    unsigned char *region;
    size_t region_stride;
    const u64 *starts;
    u64 *out;
    u32 B, n, npad, dim, k, S, R, nlo_pad;
};

template <int DIM>
__device__ __forceinline__ void load_pt(float (&p)[DIM], const float *q, u32 npad, u32 dim, u32 i) {
#pragma unroll
    for (int c = 0; c < DIM; ++c) p[c] = (c < (int)dim) ? __ldg(q + (size_t)c * npad + i) : 0.0f;
}

template <int DIM, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) kdline_async_kernel(AsyncArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const u32 C = cluster_nctarank(), rank = cluster_ctarank();
    const u32 NWK = C - 1;                       // workers
    const u32 ncl = gridDim.x / C, cl = blockIdx.x / C;
    const u32 tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    const u32 T = blockDim.x, NW = T >> 5;
    const u32 npad = a.npad, dim = a.dim, S = a.S, R = a.R;

    // ---- shared memory carve (same size in every CTA; the two roles use different parts) -----------------
    // coordinator: cand[2][32] | bound[2][32] | tab[2] | issued[16] | done[16] | res[S] | bsnd[S] | bmc[DIM][S] | pendc[R][DIM][S]
    // worker     : ring[A_RING] | bar[A_RING] | wred[32]{key,cand}
    ACand *cand = reinterpret_cast<ACand *>(smem_raw);
    u64 *bound = reinterpret_cast<u64 *>(cand + 2 * A_NC);
    ATab<DIM> *tab = reinterpret_cast<ATab<DIM> *>(bound + 2 * 32);
    u32 *issued = reinterpret_cast<u32 *>(tab + 2);
    u32 *done = issued + 16;
    u32 *rankv = done + 16;                                                   // [32] candidate ranks (atomic partial counts)
    unsigned short *ptab = reinterpret_cast<unsigned short *>(rankv + 32);   // [512] pair p -> i | j << 8
    ARes<DIM> *res = reinterpret_cast<ARes<DIM> *>(ptab + 512);
    float *bsnd = reinterpret_cast<float *>(res + S);
    float *bmc = bsnd + S;
    float *pendc = bmc + (size_t)DIM * S;
    AJob<DIM> *ring = reinterpret_cast<AJob<DIM> *>(smem_raw);
    u64 *bars = reinterpret_cast<u64 *>(ring + A_RING);
    u64 *wred_key = bars + A_RING;
    u32 *wred_cand = reinterpret_cast<u32 *>(wred_key + 32);

    // ---- one-time mailbox setup; generation counters then run across all clouds of this cluster -------------
    if (rank == 0) {
        for (u32 i = tid; i < 64; i += T) issued[i] = 0;  // issued[16] + done[16] + rankv[32]
        for (u32 p = tid; p < 496; p += T) {
            u32 j = 1;
            while ((j + 1) * j / 2 <= p) ++j;
            ptab[p] = (unsigned short)((p - j * (j - 1) / 2) | (j << 8));
        }
        for (u32 b = tid; b < S; b += T) res[b].ch[0] = make_uint4(0, 0, 0, 0);
    } else if (tid < A_RING) {
        mbar_init(smem_u32(&bars[tid]), 1);  // one arrival per job: the coordinator's remote arrive.expect_tx
    }
    if (rank != 0 && tid == 0) fence_mbar_init_cluster();
    __syncthreads();
    cluster_sync_all();

    u32 jn = 0;    // worker: jobs consumed so far
    u32 jseq = 0;  // coordinator thread: sequence number of my bucket's last job

    for (u32 cloud = cl; cloud < a.B; cloud += ncl) {
        unsigned char *rg = a.region + (size_t)cloud * a.region_stride;
        const float *q = reinterpret_cast<const float *>(rg);
        float *dis = reinterpret_cast<float *>(rg) + (size_t)dim * npad;
        const u32 *perm = reinterpret_cast<const u32 *>(dis + npad);
        const u32 *nlo = perm + npad;
        const float *fbox = reinterpret_cast<const float *>(nlo + a.nlo_pad);
        u64 *out = a.out + (size_t)cloud * a.k;

        if (rank != 0) {
            // =================================== worker ===================================================
            const u32 w = rank - 1;
            for (;; ++jn) {
                AJob<DIM> &J = ring[jn % A_RING];
                mbar_wait_cta(smem_u32(&bars[jn % A_RING]), (jn / A_RING) & 1);
                const u32 hdr = J.hdr;
                const u32 bucket = hdr & 0x7fffu;
                if (bucket == A_QUIT) {
                    __syncthreads();
                    if (tid == 0) st_cluster_u32(mapa(smem_u32(&done[w]), 0), jn + 1);
                    ++jn;
                    break;
                }
                const u32 lo = J.lo, hi = J.hi, nrefs = hdr >> 16, first = (hdr >> 15) & 1u, seq = J.seq;
                float best = -1.0f, snd = 0.0f;
                u32 bi = 0;
                float bc[DIM];
#pragma unroll
                for (int c = 0; c < DIM; ++c) bc[c] = 0.0f;
                for (u32 i = lo + tid; i < hi; i += T) {
                    float p[DIM];
                    load_pt<DIM>(p, q, npad, dim, i);
                    float v = first ? FLT_MAX : __ldcg(dis + i);
                    for (u32 r = 0; r < nrefs; ++r) {
                        float rc[DIM];
#pragma unroll
                        for (int c = 0; c < DIM; ++c) rc[c] = J.c[r * DIM + c];
                        v = fminf(v, sqdist<DIM>(p, rc));
                    }
                    __stcg(dis + i, v);
                    if (v > best) {  // ascending i: the first maximum = lowest position stays
                        snd = best < 0.0f ? 0.0f : best;
                        best = v;
                        bi = i;
#pragma unroll
                        for (int c = 0; c < DIM; ++c) bc[c] = p[c];
                    } else {
                        snd = fmaxf(snd, v);
                    }
                }
                // warp: winner key, candidate for "second largest entry"
                const u64 key = best < 0.0f ? 0ull : make_key(best, 0xfffffffeu - bi);
                const u64 wk = warp_max_key(key);
                const bool iwin = (key == wk) && key != 0ull;
                u32 cand = __float_as_uint(iwin ? snd : fmaxf(best, 0.0f));
                cand = __reduce_max_sync(FULL, cand);
                if (lane == 0) {
                    wred_key[warp] = wk;
                    wred_cand[warp] = cand;
                }
                __syncthreads();
                u64 k2 = (lane < NW) ? wred_key[lane] : 0ull;
                const u64 K = warp_max_key(k2);
                const u32 wwin = __ffs(__ballot_sync(FULL, lane < NW && k2 == K)) - 1;  // winner warp (unique position)
                u32 c2 = 0;
                if (lane < NW) c2 = (lane == wwin) ? wred_cand[lane] : (u32)(k2 >> 32);
                const u32 SND = __reduce_max_sync(FULL, c2);
                if (iwin && key == K) {  // exactly one thread: ship the result to the coordinator's mailbox
                    const u32 base = mapa(smem_u32(&res[bucket]), 0);
#pragma unroll
                    for (int ch = 1; ch < ares_chunks(DIM); ++ch) {
                        const int c0 = 3 * (ch - 1);
                        st_cluster_v4(base + 16 * ch, seq, __float_as_uint(bc[c0]),
                                      c0 + 1 < DIM ? __float_as_uint(bc[c0 + 1]) : 0u,
                                      c0 + 2 < DIM ? __float_as_uint(bc[c0 + 2]) : 0u);
                    }
                    st_cluster_v4(base, seq, (u32)(K >> 32), bi, SND);
                }
                __syncthreads();  // everyone is done with the job slot and wred
                if (tid == 0) st_cluster_u32(mapa(smem_u32(&done[w]), 0), jn + 1);
            }
        } else {
            // =================================== coordinator ===============================================
            // thread b owns bucket b: all of its state lives in registers.  The loop body is kept SMALL on purpose
            // (rolled loops, one call site per helper): warp 0's sort phase runs alone and pays every I-cache miss.
            const u32 b = tid;
            u32 st = 0;  // 0 empty / unused, 1 exact, 2 inflight
            float mx = FLT_MAX, snd = 0.0f, U = 0.0f;
            u32 pos = 0, npend = 0, blo = 0, bhi = 0;
            float lo[DIM], hi[DIM], mc[DIM];
#pragma unroll
            for (int c = 0; c < DIM; ++c) {
                lo[c] = FLT_MAX;   // empty / unused lanes do not widen the warp's box
                hi[c] = -FLT_MAX;
                mc[c] = 0.0f;
            }
            if (b < S) {
                blo = nlo[b];
                bhi = nlo[b + 1];
                if (bhi > blo) {
                    st = 1;
#pragma unroll
                    for (int c = 0; c < DIM; ++c) {
                        lo[c] = (c < (int)dim) ? fbox[(size_t)b * 2 * dim + c] : 0.0f;
                        hi[c] = (c < (int)dim) ? fbox[(size_t)b * 2 * dim + dim + c] : 0.0f;
                    }
                }
            }
            // box of the 32 buckets of this warp (consecutive leaves = one subtree, spatially tight): lets the
            // whole warp skip the per-bucket tests for a far-away pick
            float wlo[DIM], whi[DIM];
#pragma unroll
            for (int c = 0; c < DIM; ++c) {
                wlo[c] = ord2f(__reduce_min_sync(FULL, f2ord(lo[c])));
                whi[c] = ord2f(__reduce_max_sync(FULL, f2ord(hi[c])));
            }
            float *mypend = pendc + b;  // element (r, c) at mypend[(r * DIM + c) * S]
            const u32 wk_of_b = NWK ? b % NWK : 0;
            u32 first = 1;              // my next job is the bucket's first scan (KDNode::init)

            // ship my pending list as a scan job; the bucket is INFLIGHT with upper bound newU until the result lands
            auto issue = [&](float newU) {
                const u32 slot = atomicAdd(&issued[wk_of_b], 1u);
                while (slot - lds_volatile(&done[wk_of_b]) >= A_RING) {
                }
                const u32 base = mapa(smem_u32(&ring[slot % A_RING]), wk_of_b + 1);
                const u32 bar = mapa(smem_u32(&bars[slot % A_RING]), wk_of_b + 1);
                ++jseq;
                const u32 nw = npend * DIM, nch = (nw + 3) >> 2;
                mbar_arrive_expect_tx_remote(bar, 16u * (1u + nch));
                st_async_v4(base, b | (first << 15) | (npend << 16), blo, bhi, jseq, bar);
#pragma unroll 1
                for (u32 j = 0; j < nch; ++j) {
                    u32 v[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const u32 wi = 4 * j + e;
                        v[e] = wi < nw ? __float_as_uint(mypend[wi * S]) : 0u;  // (r*DIM+c)*S with r*DIM+c == wi
                    }
                    st_async_v4(base + 16 + 16 * j, v[0], v[1], v[2], v[3], bar);
                }
                st = 2;
                U = newU;
                npend = 0;
                first = 0;
            };
            // result of my in-flight job, if it arrived: take it and re-validate what was appended meanwhile
            auto poll = [&]() -> bool {
                const uint4 r0 = lds_v4_volatile(&res[b].ch[0]);
                if (r0.x != jseq) return false;
                float nc[DIM];
#pragma unroll
                for (int ch = 1; ch < ares_chunks(DIM); ++ch) {
                    const uint4 rc = lds_v4_volatile(&res[b].ch[ch]);
                    if (rc.x != jseq) return false;  // this chunk has not landed yet
                    const int c0 = 3 * (ch - 1);
                    nc[c0] = __uint_as_float(rc.y);
                    if (c0 + 1 < DIM) nc[c0 + 1] = __uint_as_float(rc.z);
                    if (c0 + 2 < DIM) nc[c0 + 2] = __uint_as_float(rc.w);
                }
                mx = __uint_as_float(r0.y);
                pos = r0.z;
                snd = __uint_as_float(r0.w);
#pragma unroll
                for (int c = 0; c < DIM; ++c) {
                    mc[c] = nc[c];
                    bmc[c * S + b] = nc[c];
                }
                bsnd[b] = snd;
                u32 keep = 0;
                float dmin = FLT_MAX;   // smallest distance of a kept reference to the new max point, if <= mx
                bool dirty = false;
#pragma unroll 1
                for (u32 r = 0; r < npend; ++r) {
                    float rc[DIM];
#pragma unroll
                    for (int c = 0; c < DIM; ++c) rc[c] = mypend[(r * DIM + c) * S];
                    if (!(boxdist<DIM>(rc, lo, hi) < mx)) continue;  // cannot lower anything in this bucket
                    const float d = sqdist<DIM>(mc, rc);
#pragma unroll
                    for (int c = 0; c < DIM; ++c) mypend[(keep * DIM + c) * S] = rc[c];
                    ++keep;
                    if (!(d > mx)) {
                        dirty = true;
                        dmin = fminf(dmin, d);
                    }
                }
                npend = keep;
                st = 1;
                // max point hit by a kept reference: rescan; survivors fill the list: flush so an append always has room
                if (dirty || npend >= R) issue(dirty ? fmaxf(snd, fminf(mx, dmin)) : mx);
                return true;
            };

            // ---- boot: the first reference = the point at POSITION start (wrapper.hpp:54-55); every leaf scans it
            //      (KDNode::init).  It runs through the same phase-C code as every later batch. -----------------------
            u32 par = 0;
            const u32 CPW = NW <= 8 ? 4u : (NW <= 16 ? 2u : 1u);   // candidates per warp, NW * CPW <= 32
            const u32 NPAIR = A_NC * (A_NC - 1) / 2;
            if (tid == 0) {
                const u32 cur = a.starts ? (u32)a.starts[cloud] : 0u;
                ATab<DIM> &Tb = tab[0];
                Tb.L0 = 1;
                Tb.bad = 0;
                Tb.pos[0] = cur;
                for (u32 c = 0; c < DIM; ++c) Tb.c[c][0] = c < dim ? __ldg(q + (size_t)c * npad + cur) : 0.0f;
            }
            __syncthreads();
            // Each iteration resolves a whole BATCH of picks.  Candidates = the CPW largest keys of every warp (32 in
            // all), sorted descending: k_0 >= k_1 >= ...  The first J of them are the next J picks of the sequential
            // recurrence, in this order, as long as for every j < J
            //   k_j is EXACT and above every bound: INFLIGHT upper bounds (they sort in front of it otherwise) and
            //       each warp's (CPW+1)-th key, i.e. every bucket that is not a candidate,
            //   no earlier pick of the batch lowers k_j's max point:        dist(P_j, P_i) >  val_j   (i < j)
            //   what is left of an earlier pick's bucket stays below k_j:   snd_i          <  val_j   (i < j)
            // (running distances only decrease, so nothing else can overtake k_j; argued in DESIGN.md).
            u64 dbg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            const bool dbg_on = (tid == 0 && cl == 0);
            bool boot = true;
            for (u32 t = 0; t < a.k;) {
                const long long c0 = clock64();
                long long c1 = c0, c2 = c0, c3 = c0;
                float wthr = FLT_MAX;  // max over this warp's buckets of (exact max | upper bound)
                if (!boot) {
                    if (st == 2) poll();
                    u64 kk = 0ull;
                    if (st == 1) kk = make_key(mx, 0xfffffffeu - pos);
                    else if (st == 2) kk = make_key(U, A_MARK);
                    // ---- phase A, every warp: its CPW largest keys (candidates) and the next one (bound) ---------
#pragma unroll 1
                    for (u32 r = 0; r <= CPW; ++r) {
                        const u64 wk = warp_max_key(kk);
                        if (r == 0) wthr = __uint_as_float((u32)(wk >> 32));
                        // the lane holding the key publishes it (EXACT keys are unique; INFLIGHT keys may repeat: the lowest of
                        // those lanes writes, and an INFLIGHT candidate only ever blocks what sorts behind it)
                        const bool me = kk == wk && wk != 0ull;
                        const u32 first_me = (u32)__ffs(__ballot_sync(FULL, me)) - 1u;
                        if (r < CPW) {
                            ACand &e = cand[par * A_NC + warp * CPW + r];
                            if (me && lane == first_me) {   // one writer: the lowest of the lanes that tie
                                e.key = wk;
                                e.bucket = warp * 32 + lane;
                            } else if (wk == 0ull && lane == 0) {
                                e.key = 0ull;
                                e.bucket = 0;
                            }
                        } else if (lane == 0) {
                            bound[par * 32 + warp] = wk;
                        }
                        if (me) kk = 0ull;
                    }
                    c1 = clock64();
                    bar_sync_id(1, T);
                    // ---- phase B0, everybody: lane l of warp w counts the candidates of warp w sorting in front of l ----
                    {
                        const u32 NCAND = NW * CPW;
                        if (lane < NCAND) {
                            const u64 myk = cand[par * A_NC + lane].key;
                            u32 cnt = 0;
#pragma unroll 1
                            for (u32 m = warp * CPW; m < warp * CPW + CPW; ++m) {
                                const u64 ok = cand[par * A_NC + m].key;
                                cnt += ((ok > myk) | ((ok == myk) & (m < lane))) ? 1u : 0u;
                            }
                            if (cnt) atomicAdd(&rankv[lane], cnt);
                        }
                    }
                    if (warp != 0) {
                        bar_arrive_id(4, T);
                    } else {
                        bar_sync_id(4, T);
                        c2 = clock64();
                        // ---- phase B1, warp 0: publish the table in sorted order ------------------------------------------
                        const bool in = lane < NW * CPW;
                        const u64 myk = in ? cand[par * A_NC + lane].key : 0ull;
                        const u32 myb = in ? cand[par * A_NC + lane].bucket : 0u;
                        const u64 bd = warp_max_key(lane < NW ? bound[par * 32 + lane] : 0ull);
                        const u32 khi = (u32)(myk >> 32), klo = (u32)myk;
                        // lanes beyond the candidates (key 0, never eligible) take the ranks behind them
                        const u32 rank = in ? rankv[lane] : lane;
                        if (in) rankv[lane] = 0;
                        ATab<DIM> &Tw = tab[par];
                        const bool elig = myk != 0ull && klo != A_MARK && myk > bd;
                        const u32 inel = __reduce_or_sync(FULL, elig ? 0u : (1u << rank));
                        Tw.pos[rank] = 0xfffffffeu - klo;
                        Tw.val[rank] = __uint_as_float(khi);
                        Tw.snd[rank] = bsnd[myb];
#pragma unroll
                        for (int c = 0; c < DIM; ++c) Tw.c[c][rank] = bmc[c * S + myb];
                        if (lane == 0) {
                            Tw.L0 = inel ? (u32)__ffs(inel) - 1u : (u32)A_NC;
                            Tw.bad = 0u;
                        }
                        c3 = clock64();
                    }
                    bar_sync_id(2, T);
                }
                ATab<DIM> &Tb = tab[par];
                par ^= 1;
                const u32 L0 = Tb.L0;
                // ---- phase B2, everybody: one candidate pair (i < j) per thread -----------------------------------
                if (L0 > 1) {
#pragma unroll 1
                    for (u32 p = tid; p < NPAIR; p += T) {
                        const u32 ij = ptab[p];
                        const u32 i = ij & 255u, j = ij >> 8;
                        if (j < L0) {
                            float Pi[DIM], Pj[DIM];
#pragma unroll
                            for (int c = 0; c < DIM; ++c) {
                                Pi[c] = Tb.c[c][i];
                                Pj[c] = Tb.c[c][j];
                            }
                            const float vj = Tb.val[j];
                            if (!(sqdist<DIM>(Pj, Pi) > vj) || !(Tb.snd[i] < vj)) atomicOr(&Tb.bad, 1u << j);
                        }
                    }
                    bar_sync_id(3, T);   // L0 is uniform, so is this barrier
                }
                const u32 badm = Tb.bad;
                u32 J = badm ? (u32)__ffs(badm) - 1u : (u32)A_NC;
                if (J > L0) J = L0;
                if (J > a.k - t) J = a.k - t;
                // ---- phase C: the J accepted picks against my bucket.  Lane j first screens pick j against the warp's
                //      box; only picks that can matter to this warp are then tested bucket by bucket, in order ----------
                bool relv = false;
                if (lane < J) {
                    float pj[DIM];
#pragma unroll
                    for (int c = 0; c < DIM; ++c) pj[c] = Tb.c[c][lane];
                    relv = boxdist<DIM>(pj, wlo, whi) < wthr;
                }
                u32 rel = __ballot_sync(FULL, relv);
                if (tid < J) out[t + tid] = Tb.pos[tid];   // positions now, original ids at the end
#pragma unroll 1
                while (rel) {
                    const u32 j = __ffs(rel) - 1;
                    rel &= rel - 1;
                    if (st == 0) continue;
                    float pc[DIM];
#pragma unroll
                    for (int c = 0; c < DIM; ++c) pc[c] = Tb.c[c][j];
                    const float bd = boxdist<DIM>(pc, lo, hi);
                    // list full while a job is out: wait for the result (rare), then test against the fresh state
                    while (st == 2 && bd < U && npend >= R) poll();
                    const float thr = st == 2 ? U : mx;
                    if (!(bd < thr)) continue;                        // cannot lower anything in this bucket
#pragma unroll
                    for (int c = 0; c < DIM; ++c) mypend[(npend * DIM + c) * S] = pc[c];
                    ++npend;
                    if (st == 2) continue;                            // job out: validated when its result lands
                    const float d = boot ? 0.0f : sqdist<DIM>(mc, pc);   // KDNode.h:122 (boot: every leaf scans)
                    if (!(d > mx))
                        issue(boot ? FLT_MAX : fmaxf(snd, fminf(mx, d)));  // max point affected: rescan
                    else if (npend >= R)
                        issue(mx);                                         // deferred list full: flush early
                }
                t += J;
                boot = false;
                if (dbg_on) {
                    const long long c4 = clock64();
                    dbg[0] += 1;
                    dbg[1] += J;
                    dbg[2] += (J == 0);
                    dbg[3] += (u64)(c1 - c0);   // poll + warp candidates
                    dbg[4] += (u64)(c2 - c1);   // warp 0 waits for the other warps
                    dbg[5] += (u64)(c3 - c2);   // sort + table
                    dbg[6] += (u64)(c4 - c3);   // pair checks + tests of the accepted picks
                }
            }
            if (dbg_on)
                for (int i = 0; i < 8; ++i) g_async_dbg[i] = dbg[i];
            // ---- drain: my job (if any) must land before the next cloud reuses the mailboxes; stop the workers ----
            while (st == 2) poll();
            __syncthreads();
            if (tid < NWK) {
                const u32 slot = atomicAdd(&issued[tid], 1u);
                while (slot - lds_volatile(&done[tid]) >= A_RING) {
                }
                const u32 base = mapa(smem_u32(&ring[slot % A_RING]), tid + 1);
                const u32 bar = mapa(smem_u32(&bars[slot % A_RING]), tid + 1);
                mbar_arrive_expect_tx_remote(bar, 16u);
                st_async_v4(base, A_QUIT, 0u, 0u, 0u, bar);
            }
            // positions -> original ids (wrapper.hpp:57-59)
            __threadfence_block();
            __syncthreads();
            for (u32 t = tid; t < a.k; t += T) {
                const u32 p = (u32)__ldcg(reinterpret_cast<const unsigned long long *>(out + t));
                out[t] = perm[p];
            }
        }
    }
    cluster_sync_all();  // nobody leaves while a peer may still write into its shared memory
}

// ======================================================================================================
//  host side
// ======================================================================================================
static int pad_dim_a(int dim) { return dim <= 2 ? 2 : dim == 3 ? 3 : dim == 4 ? 4 : dim <= 6 ? 6 : 8; }

size_t kd_region_bytes(size_t n, size_t dim, size_t h) {
    const size_t S = (size_t)1 << h, npad = (n + 31) & ~(size_t)31;
    size_t b = ((dim + 2) * npad + ((S + 1 + 31) & ~(size_t)31) + S * 2 * dim) * 4;
    return (b + 255) & ~(size_t)255;
}

template <int DIM>
static size_t async_smem(size_t S, size_t R) {
    const size_t coord = sizeof(ACand) * 2 * A_NC + 2 * 32 * 8 + sizeof(ATab<DIM>) * 2 + 32 * 4 + 32 * 4 + 512 * 2 + sizeof(ARes<DIM>) * S +
                         S * 4 + (size_t)DIM * S * 4 + R * DIM * S * 4;
    const size_t work = sizeof(AJob<DIM>) * A_RING + A_RING * 8 + 32 * 8 + 32 * 4;
    return (coord > work ? coord : work) + 16;
}

static size_t async_smem_dim(int dimp, size_t S, size_t R) {
    switch (dimp) {
        case 2: return async_smem<2>(S, R);
        case 3: return async_smem<3>(S, R);
        case 4: return async_smem<4>(S, R);
        case 6: return async_smem<6>(S, R);
        default: return async_smem<8>(S, R);
    }
}

bool plan_kdline_async(size_t n, size_t dim, size_t h, size_t B, int n_sms, AsyncPlan *pl) {
    if (dim == 0 || dim > 8 || h == 0 || h > 10 || n == 0 || B == 0) return false;
    const size_t S = (size_t)1 << h;
    const int dimp = pad_dim_a((int)dim);
    size_t R = A_MAXR;
    while (R > 2 && async_smem_dim(dimp, S, R) > 200 * 1024) --R;
    if (async_smem_dim(dimp, S, R) > 200 * 1024) return false;
    u32 threads = (u32)((S + 31) & ~(size_t)31);
    if (threads < 256) threads = 256;
    if (threads > 1024) return false;
    // cluster size: the widest cluster that still runs the whole batch in one wave; big batches trade
    // workers for clouds in flight (one worker keeps up with buckets of a few hundred points)
    const u32 Cmin = n <= 32768 ? 2 : 4;
    u32 C = 16;
    while (C > Cmin && (size_t)(n_sms / C) < B) C /= 2;
    pl->dimp = dimp;
    pl->C = C;
    pl->threads = threads;
    pl->R = (u32)R;
    pl->smem = async_smem_dim(dimp, S, R);
    size_t clusters = (size_t)n_sms / C;
    if (clusters > B) clusters = B;
    if (clusters < 1) clusters = 1;
    pl->clusters = (u32)clusters;
    return true;
}

template <int DIM, int MAXT>
static cudaError_t launch_async_tt(const AsyncPlan &pl, const AsyncArgs &a, cudaStream_t st);

template <int DIM>
static cudaError_t launch_async_t(const AsyncPlan &pl, const AsyncArgs &a, cudaStream_t st) {
    return pl.threads <= 512 ? launch_async_tt<DIM, 512>(pl, a, st) : launch_async_tt<DIM, 1024>(pl, a, st);
}

template <int DIM, int MAXT>
static cudaError_t launch_async_tt(const AsyncPlan &pl, const AsyncArgs &a, cudaStream_t st) {
    auto kern = kdline_async_kernel<DIM, MAXT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) return e;
    if (pl.C > 8) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(pl.clusters * pl.C);
    cfg.blockDim = dim3(pl.threads);
    cfg.dynamicSmemBytes = pl.smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = pl.C;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, a);
}

cudaError_t async_debug_counters(u64 *out16) {
    return cudaMemcpyFromSymbol(out16, g_async_dbg, sizeof(u64) * 16);
}

cudaError_t launch_kdline_async(const AsyncPlan &pl, unsigned char *region, size_t region_stride, const u64 *starts,
                                u64 *out, u32 B, u32 n, u32 dim, u32 k, u32 h, cudaStream_t st) {
    AsyncArgs a;
    a.region = region;
    a.region_stride = region_stride;
    a.starts = starts;
    a.out = out;
    a.B = B;
    a.n = n;
    a.npad = (n + 31) & ~31u;
    a.dim = dim;
    a.k = k;
    a.S = 1u << h;
    a.R = pl.R;
    a.nlo_pad = (a.S + 1 + 31) & ~31u;
    cudaError_t e;
    switch (pl.dimp) {
        case 2: e = launch_async_t<2>(pl, a, st); break;
        case 3: e = launch_async_t<3>(pl, a, st); break;
        case 4: e = launch_async_t<4>(pl, a, st); break;
        case 6: e = launch_async_t<6>(pl, a, st); break;
        default: e = launch_async_t<8>(pl, a, st); break;
    }
    count_launch();
    return e;
}

}  // namespace fps
